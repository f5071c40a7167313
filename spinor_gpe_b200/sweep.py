"""Batched parameter sweeps sharded across GPUs (BASELINE config 4; SURVEY.md §8e "batched sweep").

Independent trajectories never interact, so there is no data-path collective: trajectory ``i`` runs on rank
``i mod world`` (one process per GPU), ``batch`` trajectories at a time in one plan (``[B][2][Ny][Nx]`` state,
per-trajectory potentials via a batch stride, per-trajectory uniform couplings), and only the small results
(populations, energies, optionally the final states) are gathered at the end.

The reference has no batch concept (one ``PSpinor`` per run); the oracle for a sweep is a Python loop over
single runs, which is what ``tests/test_gpu_parity.py::test_batched_sweep...`` and the gloo test compare with.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _capi
from ._separable import split_separable


def shard(n_items, rank, world):
    """Indices owned by ``rank``: i mod world == rank (round-robin keeps shards balanced to within one)."""
    return list(range(rank, n_items, world))


class Trajectory:
    """Per-trajectory overrides of a base ``PSpinor``: uniform coupling Omega and the two potential grids
    (``pot_eng_spin``, i.e. trap +- detuning/2); ``psik`` defaults to the base state."""

    def __init__(self, omega, pot_eng_spin, psik=None, label=None):
        self.omega = float(omega)
        self.pot = np.asarray(pot_eng_spin, dtype=np.float64)
        self.psik = None if psik is None else np.asarray(psik)
        self.label = label


def detuning_coupling_grid(ps, couplings, slopes, axis=1):
    """The 8x8 style sweep of examples/4_detuning_grad.py: every (coupling, detuning-gradient) pair."""
    mesh = ps.space['x_mesh'] if axis == 0 else ps.space['y_mesh']
    out = []
    for om in couplings:
        for sl in slopes:
            det = mesh * sl
            out.append(Trajectory(om, [ps.pot_eng + det / 2, ps.pot_eng - det / 2], label=(float(om), float(sl))))
    return out


def _default_factory(nx, ny, batch, dtype, device):
    from .plan import Plan
    return Plan(nx, ny, batch, dtype, device)


def run_sweep(ps, trajectories, t_step, n_steps, time='imag', device='cuda', batch=8, precision='c128',
              keep_states=False, group=None, plan_factory=None, energy=True, unwrap='none'):
    """Propagate every trajectory for ``n_steps`` full steps; returns on every rank the gathered dict
    ``{'pops': (n, n_steps, 2), 'energy': (n, 4) or None, 'psik': list or None, 'owner': (n,)}``.

    ``unwrap``: phase treatment of the final energies — 'none' (default: stays on the device, asynchronous),
    'local', or 'herraez' (the reference's unwrapped phase; the region merging of every trajectory runs on the host).

    ``ps`` supplies the shared grid, kinetic energy, interactions, atom number and (default) initial state.
    Works with any initialised ``torch.distributed`` group (NCCL on GPUs; gloo in the CPU tests, where
    ``plan_factory`` substitutes the emulated plan) or without one (single process)."""
    use_dist = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if use_dist else 0
    world = dist.get_world_size(group) if use_dist else 1
    factory = plan_factory or _default_factory
    cdtype = torch.complex128 if precision == 'c128' else torch.complex64
    n = len(trajectories)
    mine = shard(n, rank, world)
    ny, nx = np.asarray(ps.psik[0]).shape
    kin = np.array([np.asarray(k) for k in ps.kin_eng_spin])
    ksep = split_separable(kin)
    if not ps.rot_coupling and ps.is_coupling:
        eiphi = np.exp(1j * 2 * ps.kL_recoil * np.asarray(ps.space['x']))
    else:
        eiphi = None
    base_psik = np.array([np.asarray(p) for p in ps.psik])

    local = {}
    for start in range(0, len(mine), batch):
        ids = mine[start:start + batch]
        B = len(ids)
        pl = factory(nx, ny, B, cdtype, device)
        pl.set_grid(ps.space['dr'][0], ps.space['dr'][1], ps.space['dv_r'], ps.space['dv_k'], ps.atom_num)
        pl.set_interactions(ps.g_sc['uu'], ps.g_sc['dd'], ps.g_sc['ud'])
        pl.set_kinetic(kin[0], kin[1])
        if ksep is not None:
            pl.set_kinetic_separable(*ksep)
        pots = np.stack([trajectories[i].pot for i in ids])                 # (B, 2, ny, nx)
        pl.set_potential(np.ascontiguousarray(pots[:, 0]), np.ascontiguousarray(pots[:, 1]), batched=True)
        seps = [split_separable(p) for p in pots]
        if all(s is not None for s in seps):
            pl.set_potential_separable(np.stack([s[0] for s in seps]), np.stack([s[1] for s in seps]), batched=True)
        if ps.is_coupling:
            pl.set_coupling(_capi.SGPE_COUPLING_UNIFORM, omega=np.array([trajectories[i].omega for i in ids]),
                            eiphi=eiphi)
        else:
            pl.set_coupling(_capi.SGPE_COUPLING_NONE)
            omegas = np.array([trajectories[i].omega for i in ids], dtype=np.float64)
            if np.any(omegas):      # eng_expect adds the coupling energy whatever is_coupling says (:319-321)
                pl.set_energy_coupling(_capi.SGPE_COUPLING_UNIFORM, omega=omegas)
        pl.set_time(time, t_step)
        states = np.stack([base_psik if trajectories[i].psik is None else trajectories[i].psik for i in ids])
        pl.load(states)
        pops = torch.zeros((B, max(n_steps, 1), 2), dtype=torch.float64, device=device)
        pl.full_steps(n_steps, pops)
        en = pl.energy(None, kl_term=2 * ps.kL_recoil * float(bool(ps.is_coupling)), unwrap=unwrap) if energy else None
        final = pl.store() if keep_states else None
        pops_h = pops.cpu().numpy()[:, :n_steps]
        for k, i in enumerate(ids):
            local[i] = (pops_h[k].copy(), None if en is None else en[k].cpu().numpy().copy(),
                        None if final is None else final[k].cpu().numpy().copy())
        pl.close()

    if use_dist:
        gathered = [None] * world
        dist.all_gather_object(gathered, local, group=group)
    else:
        gathered = [local]
    merged, owner = {}, np.zeros(n, dtype=np.int64)
    for r, part in enumerate(gathered):
        for i, val in part.items():
            merged[i] = val
            owner[i] = r
    assert sorted(merged) == list(range(n)), "sweep shards do not cover every trajectory exactly once"
    return {
        'pops': np.stack([merged[i][0] for i in range(n)]),
        'energy': np.stack([merged[i][1] for i in range(n)]) if energy else None,
        'psik': [merged[i][2] for i in range(n)] if keep_states else None,
        'owner': owner,
    }
