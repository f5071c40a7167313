"""Slab-decomposed propagation of ONE grid over several GPUs (BASELINE config 5; SURVEY.md §8e "slab FFT").

Rank r of P owns the rows y in [r Ny/P, (r+1) Ny/P).  A split sub-step is

    k-junction on the TRANSPOSED slab  [2][Nx/P][Ny]  (contiguous y-lines): FFT_y, K factors, norm sums, iFFT_y
    all-reduce of the three norm sums (T, S0, S1)
    all-to-all transpose  -> row slab [2][Ny/P][Nx]
    row pass: iFFT_x, normalise with the GLOBAL sum, I C P C I, FFT_x
    all-to-all transpose  -> transposed slab

i.e. two all-to-alls per sub-step and k-space kept in the transposed layout, so a 2-D transform costs one
exchange.  The local passes are the hand-written kernels behind ``sgpe_pass_rows / sgpe_pass_klines /
sgpe_slab_pack / sgpe_slab_unpack``; ``torch.distributed`` (NCCL over NVLink) carries the collectives.
Line lengths are limited to 4096 points this round, so this path is exercised on grids that
would also fit one GPU and is checked against the single-GPU propagator.
"""
import math

import numpy as np
import torch
import torch.distributed as dist

from . import _capi
from ._separable import split_separable

MAGIC_GAMMA = 1 / (2 + 2 ** (1 / 3))


class SlabPropagator:
    """Distributed TensorPropagator for one trajectory.  Every rank passes the same (global) ``PSpinor``; each
    keeps only its slab on the device."""

    def __init__(self, spin, t_step, time='imag', device='cuda', group=None, precision='c128', plan_kwargs=None):
        from .plan import Plan
        assert dist.is_initialized(), "SlabPropagator needs an initialised torch.distributed process group"
        self.group = group
        self.rank, self.P = dist.get_rank(group), dist.get_world_size(group)
        self.dev = torch.device(device)
        self.cdtype = torch.complex128 if precision == 'c128' else torch.complex64
        psik = np.array([np.asarray(p) for p in spin.psik])
        _, self.ny, self.nx = psik.shape
        P, r = self.P, self.rank
        assert self.nx % (32 * P) == 0 and self.ny % (32 * P) == 0, "mesh must split into multiples of 32 per rank"
        self.nxl, self.nyl = self.nx // P, self.ny // P
        self.time, self.dt = time, float(t_step)
        self.dt_out, self.dt_in = self.dt * MAGIC_GAMMA, self.dt * (1 - 2 * MAGIC_GAMMA)
        self.atom_num = float(spin.atom_num)
        self.dv_k = float(spin.space['dv_k'])
        kw = dict(plan_kwargs or {})
        ys, xs = slice(r * self.nyl, (r + 1) * self.nyl), slice(r * self.nxl, (r + 1) * self.nxl)

        # ---- row plan: local rows, full x
        self.rp = rp = Plan(self.nx, self.nyl, 1, self.cdtype, self.dev, **kw)
        dr = spin.space['dr']
        rp.set_grid(dr[0], dr[1], spin.space['dv_r'], spin.space['dv_k'], spin.atom_num)
        rp.set_interactions(spin.g_sc['uu'], spin.g_sc['dd'], spin.g_sc['ud'])
        pot = np.array([np.asarray(v) for v in spin.pot_eng_spin])
        rp.set_potential(np.ascontiguousarray(pot[0, ys]), np.ascontiguousarray(pot[1, ys]))
        psep = split_separable(pot)
        if psep is not None:
            rp.set_potential_separable(psep[0], np.ascontiguousarray(psep[1][:, ys]))
        cpl = np.asarray(spin.coupling, dtype=np.float64)
        eiphi = None
        if spin.is_coupling and not spin.rot_coupling:
            eiphi = np.exp(1j * 2 * spin.kL_recoil * np.asarray(spin.space['x']))
        if not spin.is_coupling or not np.any(cpl):
            rp.set_coupling(_capi.SGPE_COUPLING_NONE)
        elif np.all(cpl == cpl.flat[0]):
            rp.set_coupling(_capi.SGPE_COUPLING_UNIFORM, omega=np.array([cpl.flat[0]]), eiphi=eiphi)
        else:
            rp.set_coupling(_capi.SGPE_COUPLING_DENSE, coupling=np.ascontiguousarray(cpl[ys]), eiphi=eiphi)
        rp.set_time(time, self.dt)

        # ---- transposed plan: lines = local k_x, positions = k_y
        self.tp = tp = Plan(self.ny, self.nxl, 1, self.cdtype, self.dev, **kw)
        tp.set_grid(dr[0], dr[1], spin.space['dv_r'], spin.space['dv_k'], spin.atom_num)
        kin = np.array([np.asarray(k) for k in spin.kin_eng_spin])
        kin_t = np.ascontiguousarray(kin[:, :, xs].transpose(0, 2, 1))          # (2, nxl, ny)
        tp.set_kinetic(kin_t[0], kin_t[1])
        ksep = split_separable(kin)
        if ksep is not None:      # kin[ky][kx] = gx[kx] + gy[ky]:  position table <- gy, line table <- local gx
            tp.set_kinetic_separable(ksep[1], np.ascontiguousarray(ksep[0][:, xs]))
        tp.set_time(time, self.dt)

        n_local = 2 * self.nxl * self.ny
        mk = lambda: torch.empty(n_local, dtype=self.cdtype, device=self.dev)      # noqa: E731
        self.tbuf, self.rbuf, self.send, self.recv = mk(), mk(), mk(), mk()
        self.sums = torch.zeros(4, dtype=torch.float64, device=self.dev)
        local = np.ascontiguousarray(psik[:, :, xs].transpose(0, 2, 1))          # (2, nxl, ny) transposed slab
        self.tbuf.copy_(torch.as_tensor(local).reshape(-1).to(self.cdtype))
        self.mid = False
        self.pending_dt = 0.0
        self.scale_pending = False
        self.a2a_bytes = 0

    # ------------------------------------------------------------------ collectives
    def _all_to_all(self):
        s, r = torch.view_as_real(self.send), torch.view_as_real(self.recv)
        dist.all_to_all_single(r.view(-1), s.view(-1), group=self.group)
        self.a2a_bytes += self.send.numel() * self.send.element_size() * (self.P - 1) // self.P

    def _to_rows(self):
        self.tp.slab_pack(self.tbuf, self.send, self.nxl, self.P, self.nyl)      # [2][nxl][P*nyl] -> [P][2][nxl][nyl]
        self._all_to_all()
        self.rp.slab_unpack(self.recv, self.rbuf, self.P, self.nxl, self.nyl)    # -> [2][nyl][P*nxl]

    def _to_lines(self):
        self.rp.slab_pack(self.rbuf, self.send, self.nyl, self.P, self.nxl)      # [2][nyl][P*nxl] -> [P][2][nyl][nxl]
        self._all_to_all()
        self.tp.slab_unpack(self.recv, self.tbuf, self.P, self.nyl, self.nxl)    # -> [2][nxl][P*nyl]

    def _reduce_sums(self):
        dist.all_reduce(self.sums, op=dist.ReduceOp.SUM, group=self.group)

    # ------------------------------------------------------------------ stepping
    def single_step(self, dt_sub, pops_out=None):
        imag = (self.time == 'imag')
        if not self.mid:
            self.tp.pass_klines(self.tbuf, False, False, 0.0, True, dt_sub / 2, True, self.sums)
        elif pops_out is not None and imag:
            self.tp.pass_klines(self.tbuf, True, True, self.pending_dt / 2, True, dt_sub / 2, True, self.sums)
        else:
            self.tp.pass_klines(self.tbuf, True, False, 0.0, True, (self.pending_dt + dt_sub) / 2, True, self.sums)
        self._reduce_sums()
        if pops_out is not None and self.mid:
            pops_out.copy_(self.atom_num * self.sums[1:3] / (self.sums[1] + self.sums[2]))
        self._to_rows()
        self.rp.pass_rows(self.rbuf, dt_sub, self.sums, float(self.nx) * float(self.ny))
        self._to_lines()
        self.mid, self.pending_dt, self.scale_pending = True, dt_sub, False

    def close_junction(self, pops_out=None):
        if not self.mid:
            return
        self.tp.pass_klines(self.tbuf, True, True, self.pending_dt / 2, False, 0.0, False, self.sums)
        self._reduce_sums()
        if pops_out is not None:
            pops_out.copy_(self.atom_num * self.sums[1:3] / (self.sums[1] + self.sums[2]))
        self.mid, self.scale_pending = False, True

    def full_steps(self, n, pops=None):
        """n full steps; pops: optional (n, 2) float64 tensor on the device (same on every rank)."""
        pending = None
        for i in range(n):
            self.single_step(self.dt_out, pending)
            self.single_step(self.dt_in)
            self.single_step(self.dt_out)
            pending = pops[i] if pops is not None else None
        if n > 0:
            self.close_junction(pending)

    def local_psik(self):
        """Normalised k-space state of this rank in the transposed layout, (2, Nx/P, Ny)."""
        self.close_junction()
        out = self.tbuf.view(2, self.nxl, self.ny)
        if self.scale_pending:
            scale = torch.sqrt(self.atom_num / (self.dv_k * (self.sums[1] + self.sums[2])))
            out = out * scale.to(out.real.dtype)
        return out

    def gather_psik(self):
        """Full (2, Ny, Nx) k-space state on every rank (tests / small grids only)."""
        loc = torch.view_as_real(self.local_psik().contiguous())
        parts = [torch.empty_like(loc) for _ in range(self.P)]
        dist.all_gather(parts, loc, group=self.group)
        full = torch.cat([torch.view_as_complex(p) for p in parts], dim=1)       # (2, Nx, Ny)
        return full.transpose(1, 2).contiguous()
