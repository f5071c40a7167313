#!/usr/bin/env python
"""bench.py — full split-steps per second of the B200 propagator on BASELINE.json's headline config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mesh 2048]

Headline workload (BASELINE.json configs[2], the set-up of the reference's benchmarks/benchmark_prop.py:50-67):
imaginary-time ground state, 2048x2048 complex128, atom_num=1e2, omega=2pi*(50,50,2000), g=(1,1,1.04),
r_sizes=(8,8), coupling_setup(kin_shift=False) with Omega=0, dt=1/50, per-step renormalisation and per-step
populations; one "step" = one TensorPropagator.full_step() (3 split sub-steps).

The ONE JSON line rank 0 prints carries, next to the contract's keys (value, e2e, roofline, cpu_baseline, clocks ...):

    energy_tracking  the headline workload with eng_expect after every step (the "energy tracking" of configs[2])
    variants         the same mesh through the general dense-operator kernels, in real time, and in complex64
    configs          configs[0] (256^2 ground state) and configs[1] (1024^2 real-time Raman with momentum kick)
    library_bar      the reference's own op sequence (torch ATen + cuFFT) on the same GPU (benchmark_prop.py method)
    e2e_public       the same run through PSpinor.imaginary() — the call a user of the reference makes
    sweep            configs[3]: 64 x 512^2 trajectories, strong-scaled over the N ranks (no data-path collective)
    slab             (N >= 2) configs[4]: one mesh row-sharded over the N ranks with the fused NVLink exchange, and
                     slab_parity = slab (p2p and nccl exchange, forced four-step lines) vs the single GPU; a parity
                     above 1e-10 makes the process exit non-zero after the line is printed

N > 1 (launched by torch.distributed.run, one rank per GPU): the headline is one independent 2048^2 trajectory per
rank (a sweep over g_ud), no data-path collective -> "scaling": "weak"; the sweep / slab records are the two
multi-GPU modes of the north star.  See DESIGN.md "Measurement" for what each key means.
"""
import argparse
import json
import os
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W0 = 2 * np.pi * 50
MODE = 'imag'                       # --mode real switches the headline run to real-time propagation
DT = {'imag': 1 / 50, 'real': 1 / 5000}
NVLINK_NOMINAL, NVLINK_MEASURED = 900.0, 770.0      # GB/s per direction per GPU (B200_PROFILING.md)
TOL_PARITY = 1e-10


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def build_problem(mesh, g_ud=1.04, tag='bench'):
    """PSpinor set-up of benchmark_prop.py (host NumPy, one-off)."""
    from spinor_gpe_b200 import PSpinor
    tmp = os.path.join(tempfile.mkdtemp(prefix='sgpe_bench_'), tag) + os.sep
    ps = PSpinor(tmp, overwrite=True, atom_num=1e2, omeg={'x': W0, 'y': W0, 'z': 40 * W0},
                 g_sc={'uu': 1, 'dd': 1, 'ud': g_ud}, pop_frac=(0.5, 0.5), r_sizes=(8, 8),
                 mesh_points=(mesh, mesh))
    ps.coupling_setup(wavel=790.1e-9, kin_shift=False)
    return ps


def build_raman_problem(mesh, tag='raman'):
    """BASELINE configs[1]: examples/3_raman_rabi.py:126-183 at mesh^2 (Raman set-up with the spin-dependent kinetic
    shift, momentum kick, laboratory-frame coupling phase, uniform coupling of one recoil energy)."""
    from spinor_gpe_b200 import PSpinor
    tmp = os.path.join(tempfile.mkdtemp(prefix='sgpe_bench_'), tag) + os.sep
    ps = PSpinor(tmp, overwrite=True, atom_num=1e4, omeg={'x': W0, 'y': W0, 'z': 40 * W0},
                 g_sc={'uu': 1, 'dd': 1, 'ud': 0.0}, pop_frac=(1.0, 0.0), r_sizes=(16, 16), mesh_points=(mesh, mesh))
    ps.coupling_setup(wavel=790.1e-9, kin_shift=True)
    ps.shift_momentum(scale=1.0, frac=(0, 1.0))
    ps.rot_coupling = False
    ps.coupling_uniform(1.0 * ps.EL_recoil)
    return ps


def hbm_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def recorded_traffic(mesh):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            return json.load(f).get(str(mesh))
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: 'hw_slowdown',
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: 'hw_thermal_slowdown',
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: 'sw_thermal_slowdown',
                 nv.nvmlClocksThrottleReasonSwPowerCap: 'sw_power_cap'}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def summary(self):
        return {'sm_mhz': float(np.median(self.sm)) if self.sm else None, 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(self.sm)}


# ----------------------------------------------------------------------------- the reference's CPU path
def oracle_problem(ps):
    from oracle import spinor_oracle as orc
    return orc.Problem(ps.psik, ps.kin_eng_spin, ps.pot_eng_spin, ps.coupling, ps.space['dr'],
                       ps.space['dv_r'], ps.space['dv_k'], [ps.g_sc['uu'], ps.g_sc['dd'], ps.g_sc['ud']],
                       ps.atom_num, x=ps.space['x'], kL=ps.kL_recoil, is_coupling=ps.is_coupling,
                       rot_coupling=ps.rot_coupling)


def oracle_steps_per_s(ps, max_steps, warmup, budget_s=40.0):
    """The CPU arm: the oracle's torch-CPU restatement of TensorPropagator.full_step on ALL host cores.  Under
    torch.distributed.run OMP_NUM_THREADS defaults to 1: the thread count is set explicitly."""
    import torch
    from oracle import spinor_oracle as orc
    torch.set_num_threads(host_cores())
    prob = oracle_problem(ps)
    o = orc.OraclePropagator(prob, DT[MODE], MODE)
    t0 = time.perf_counter()
    o.full_step()
    t_first = time.perf_counter() - t0
    for _ in range(max(0, warmup - 1)):
        o.full_step()
    n = int(max(3, min(max_steps, budget_s / max(t_first, 1e-3))))
    times = []
    for _ in range(n):
        t0 = time.perf_counter()
        o.full_step()
        orc.populations(o.psik, prob.dv_k)           # prop_loop does this every step (:194)
        times.append(time.perf_counter() - t0)
    med = float(np.median(times))
    return 1.0 / med, n, torch.get_num_threads(), med


def library_bar_steps_per_s(ps, dev, n_steps=10):
    """The reference's own op sequence (torch ATen element-wise + torch.fft = cuFFT, two .item() syncs per
    single_step) on the SAME GPU: the oracle's restatement with its tensors moved to the device, timed the way
    benchmarks/benchmark_prop.py:58-97 times prop.full_step().  A baseline leg like cpu_baseline; never on the
    product path."""
    import torch
    from oracle import spinor_oracle as orc
    prob = oracle_problem(ps)
    for name in ('psik', 'kin', 'pot', 'coupling', 'expon'):
        setattr(prob, name, getattr(prob, name).to(dev))
    o = orc.OraclePropagator(prob, DT[MODE], MODE)
    for _ in range(3):
        o.full_step()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(n_steps):
        o.full_step()
        orc.populations(o.psik, prob.dv_k)
    torch.cuda.synchronize(dev)
    return n_steps / (time.perf_counter() - t0)


def workload_config(args):
    """Identical in both arms (the driver compares the two `config` dicts)."""
    kind = 'imaginary-time ground state' if MODE == 'imag' else 'real-time propagation (dt=1/5000)'
    return {'workload': f'{kind} {args.mesh}x{args.mesh} {args.precision}, benchmark_prop.py '
                        'parameters (atom_num=1e2, g=(1,1,1.04), r_sizes=(8,8), coupling_setup(kin_shift=False), '
                        'dt=1/50), per-step renormalisation + populations (BASELINE configs[2])',
            'mesh': [args.mesh, args.mesh], 'trajectories_per_gpu': 1,
            'l2': 'state (128 MiB at 2048^2) exceeds the 126 MB L2; no explicit flush'}


def run_reference(args, rank, world):
    if rank != 0:
        return
    ps = build_problem(args.mesh)
    sps, n, cores, med = oracle_steps_per_s(ps, args.steps, args.warmup, budget_s=60.0)
    sample = f'{n} full_step() calls of the {args.mesh}x{args.mesh} workload after warm-up, median'
    line = {
        'impl': 'reference', 'metric': 'full split-steps/s', 'value': sps, 'unit': 'steps/s', 'n_gpus': 0,
        'steps': n, 'warmup': args.warmup, 'ms_per_step': med * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args), 'where': f'host cores of the GPU box ({cores} threads)',
        'cpu_baseline': {'value': sps, 'unit': 'steps/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': sps, 'unit': 'steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- helpers of the GPU legs
class Ctx:
    """What every leg needs: rank / world, device, barrier, max-over-ranks."""

    def __init__(self, rank, world, local_rank):
        import torch
        self.rank, self.world, self.local = rank, world, local_rank
        self.dev = torch.device('cuda', local_rank)

    def barrier(self):
        import torch
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, ms):
        import torch
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def timed(self, fn):
        """fn() bracketed by barrier + synchronize and CUDA events on the launching stream; max over ranks, ms."""
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1))


def plan_for(ps, dev, precision='c128', mode='imag', dt=None, dense=False, options=None):
    """A device-resident plan of a PSpinor problem (what TensorPropagator.__init__ sets up), state loaded."""
    import torch
    from spinor_gpe_b200 import _capi
    from spinor_gpe_b200._separable import split_separable
    from spinor_gpe_b200.plan import Plan
    cdtype = torch.complex128 if precision == 'c128' else torch.complex64
    ny, nx = np.asarray(ps.psik[0]).shape
    pl = Plan(nx, ny, 1, cdtype, dev)
    pl.set_grid(ps.space['dr'][0], ps.space['dr'][1], ps.space['dv_r'], ps.space['dv_k'], ps.atom_num)
    pl.set_interactions(ps.g_sc['uu'], ps.g_sc['dd'], ps.g_sc['ud'])
    pl.set_time(mode, DT[mode] if dt is None else dt)
    kin, pot = np.array(ps.kin_eng_spin), np.array(ps.pot_eng_spin)
    ksep = None if dense else split_separable(kin)
    psep = None if dense else split_separable(pot)
    if ksep is not None:
        pl.set_kinetic_separable(*ksep)
    else:
        pl.set_kinetic(kin[0], kin[1])
    if psep is not None:
        pl.set_potential_separable(*psep)
    else:
        pl.set_potential(pot[0], pot[1], shared=bool(np.array_equal(pot[0], pot[1])))
    cpl = np.asarray(ps.coupling, dtype=np.float64)
    if not ps.is_coupling or not np.any(cpl):
        pl.set_coupling(_capi.SGPE_COUPLING_NONE)        # Omega == 0: C is exactly the identity
    else:
        eiphi = None if ps.rot_coupling else np.exp(1j * 2 * ps.kL_recoil * np.asarray(ps.space['x']))
        assert np.all(cpl == cpl.flat[0])
        pl.set_coupling(_capi.SGPE_COUPLING_UNIFORM, omega=np.array([cpl.flat[0]]), eiphi=eiphi)
    for k, v in (options or {}).items():
        pl.set_option(k, v)
    pl.load(np.array(ps.psik)[None])
    pl.operators = ('separable factor tables (auto-detected)' if ksep is not None and psep is not None
                    else 'dense grids, factors evaluated per point')
    return pl


def steps_per_s(ctx, pl, steps, warmup, track_launches=False):
    import torch
    pops = torch.zeros((1, steps + warmup, 2), dtype=torch.float64, device=ctx.dev)
    pl.full_steps(warmup, pops, first=0)
    l0 = pl.launch_count()
    ms = ctx.timed(lambda: pl.full_steps(steps, pops, first=warmup))
    launches = pl.launch_count() - l0
    out = {'value': 1e3 * steps / ms, 'unit': 'steps/s', 'ms_per_step': ms / steps, 'steps': steps}
    if track_launches:
        out['launches_per_step'] = launches / steps
    return out, pops


def frac_of_roofline(sps, mesh_points, precision='c128', batch=1):
    bytes_per_step = (768.0 if precision == 'c128' else 384.0) * mesh_points * batch
    return bytes_per_step * sps / 1e9 / hbm_peak()[0]


def leg(line, key, fn):
    """Run an optional leg; a failure is reported under its key and never hides the headline numbers."""
    try:
        line[key] = fn()
    except Exception as exc:                         # noqa: BLE001
        line[key] = {'error': f'{type(exc).__name__}: {exc}'}


# ----------------------------------------------------------------------------- multi-GPU legs
def sweep_leg(ctx, steps=30, warm=5):
    """BASELINE configs[3]: 64 independent 512x512 trajectories (8 couplings x 8 detuning gradients of
    examples/4_detuning_grad.py:118-157), imaginary time, trajectory i -> rank i mod N, all trajectories of a rank in ONE
    batched plan; no data-path collective.  Strong scaling: the 64 trajectories are fixed, N varies."""
    import torch
    from spinor_gpe_b200 import PSpinor, _capi
    from spinor_gpe_b200._separable import split_separable
    from spinor_gpe_b200.plan import Plan
    from spinor_gpe_b200.sweep import detuning_coupling_grid, shard
    mesh = 512
    ps = PSpinor(os.path.join(tempfile.mkdtemp(prefix='sgpe_sweep_'), f'r{ctx.rank}') + os.sep, overwrite=True,
                 atom_num=1e4, omeg={'x': W0, 'y': W0, 'z': 40 * W0}, g_sc={'uu': 1, 'dd': 0.995, 'ud': 0.995},
                 pop_frac=(0.5, 0.5), r_sizes=(16, 16), mesh_points=(mesh, mesh))
    ps.coupling_setup(wavel=804e-9, kin_shift=True)
    ps.shift_momentum(scale=0.6, frac=(0.5, 0.5))
    trajs = detuning_coupling_grid(ps, np.linspace(0.5, 5, 8) * ps.EL_recoil, np.linspace(-12, 12, 8))
    mine = shard(len(trajs), ctx.rank, ctx.world)
    B = len(mine)
    pl = Plan(mesh, mesh, B, torch.complex128, ctx.dev)
    pl.set_grid(ps.space['dr'][0], ps.space['dr'][1], ps.space['dv_r'], ps.space['dv_k'], ps.atom_num)
    pl.set_interactions(ps.g_sc['uu'], ps.g_sc['dd'], ps.g_sc['ud'])
    pl.set_kinetic_separable(*split_separable(np.array(ps.kin_eng_spin)))
    seps = [split_separable(trajs[i].pot) for i in mine]
    pl.set_potential_separable(np.stack([s[0] for s in seps]), np.stack([s[1] for s in seps]), batched=True)
    pl.set_coupling(_capi.SGPE_COUPLING_UNIFORM, omega=np.array([trajs[i].omega for i in mine]))
    pl.set_time('imag', 1 / 50)
    pl.load(np.stack([np.array(ps.psik)] * B))
    pops = torch.zeros((B, steps + warm, 2), dtype=torch.float64, device=ctx.dev)
    pl.full_steps(warm, pops)
    ms = ctx.timed(lambda: pl.full_steps(steps, pops, first=warm)) / steps
    tsps = len(trajs) * 1e3 / ms
    per_gpu = 768.0 * mesh * mesh * B / (ms * 1e-3) / 1e9
    atoms = float(pops[0, -1].sum())
    pl.close()
    return {'workload': '64 trajectories x 512^2 c128, imaginary time, 8 couplings x 8 detuning gradients '
                        '(BASELINE configs[3]); strong scaling over the ranks, no data-path collective',
            'value': tsps, 'unit': 'trajectory-steps/s', 'ranks': ctx.world, 'trajectories': len(trajs),
            'per_rank_batch': B, 'ms_per_sweep_step': ms, 'steps': steps,
            'hbm_algorithmic_GBps_per_gpu': per_gpu, 'hbm_frac_per_gpu': per_gpu / hbm_peak()[0],
            'atom_number_check': atoms}


def slab_parity_leg(ctx, mesh=4096):
    """Slab propagation (both exchange variants, forced four-step lines) against the single-GPU propagator on the
    same problem: psi_k rel-L2 and populations after two full steps.  Rank 0 holds the single-GPU result."""
    import torch
    from spinor_gpe_b200 import TensorPropagator
    from spinor_gpe_b200.slab import SlabPropagator
    ps = build_problem(mesh, tag=f'slabpar{ctx.rank}')
    ps.coupling_uniform(0.5 * ps.EL_recoil)            # a non-trivial coupling operator
    ps.rot_coupling = False
    out, worst = [], 0.0
    n = 2
    for mode, dt in (('real', 1 / 5000), ('imag', 1 / 50)):
        ref = pops1 = None
        if ctx.rank == 0:
            prop = TensorPropagator(ps, dt, n, ctx.dev, time=mode)
            pops1 = torch.zeros((1, n, 2), dtype=torch.float64, device=ctx.dev)
            prop._plan.full_steps(n, pops1)
            ref = torch.stack(prop.psik)
            del prop
        for exchange, splits in (('p2p', (64, 64)), ('nccl', (64, 32)), ('p2p', (None, None))):
            sp = SlabPropagator(ps, dt, time=mode, device=ctx.dev, split_x=splits[0], split_y=splits[1],
                                exchange=exchange)
            pops = torch.zeros((n, 2), dtype=torch.float64, device=ctx.dev)
            sp.full_steps(n, pops)
            full = sp.gather_psik()
            rec = {'mesh': mesh, 'mode': mode, 'exchange': exchange, 'four_step_split': list(splits)}
            if ctx.rank == 0:
                rec['psik_rel_l2'] = float(torch.linalg.norm(full - ref) / torch.linalg.norm(ref))
                rec['pops_rel'] = float((pops - pops1[0]).abs().max() / pops1.abs().max())
                worst = max(worst, rec['psik_rel_l2'], rec['pops_rel'] / 10)    # pops tolerance 1e-9
                out.append(rec)
            sp.close()
            del sp, full
        del ref
        torch.cuda.empty_cache()
    return {'against': 'single-GPU TensorPropagator on rank 0', 'steps': n, 'ranks': ctx.world, 'cases': out,
            'worst_psik_rel_l2': max([c['psik_rel_l2'] for c in out], default=None),
            'tolerance': TOL_PARITY, 'ok': bool(worst <= TOL_PARITY)}


def slab_leg(ctx, steps=10):
    """BASELINE configs[4]: ONE mesh row-sharded over the ranks, real time with uniform Raman coupling, set up from
    1-D vectors (no (Ny, Nx) host array); the last pass of each direction stores straight into the peers' buffers
    (CUDA IPC over NVLink).  16384^2 at 8 ranks (the config as written), 8192^2 at 2 / 4."""
    import torch
    from spinor_gpe_b200.slab import SeparableProblem, SlabPropagator
    mesh = 16384 if ctx.world >= 8 else 8192
    prob = SeparableProblem((mesh, mesh), r_sizes=(64, 64), atom_num=1e6, omeg={'x': W0, 'y': W0, 'z': 40 * W0},
                            g_sc={'uu': 1, 'dd': 1, 'ud': 1.04}, pop_frac=(0.5, 0.5), coupling=1.0, kin_shift=True,
                            rot_coupling=False)
    sp = SlabPropagator(prob, 1 / 5000, time='real', device=ctx.dev, exchange='auto')
    pops = torch.zeros((steps, 2), dtype=torch.float64, device=ctx.dev)
    sp.full_steps(2)
    ctx.barrier()
    sp.a2a_bytes = 0
    l0 = sp.rp.launch_count() + sp.tp.launch_count()
    ms = ctx.timed(lambda: sp.full_steps(steps, pops)) / steps
    sent = sp.a2a_bytes / steps
    nvl = sent / (ms * 1e-3) / 1e9
    hbm = 768.0 * mesh * mesh / ctx.world / (ms * 1e-3) / 1e9
    rec = {'workload': f'real-time propagation {mesh}x{mesh} c128, uniform Raman coupling, row-sharded over '
                       f'{ctx.world} ranks (BASELINE configs[4])',
           'value': 1e3 / ms, 'unit': 'steps/s', 'ms_per_step': ms, 'steps': steps, 'ranks': ctx.world,
           'mesh': [mesh, mesh], 'exchange': sp.exchange, 'four_step_split': [sp.n1x, sp.n1y],
           'windows': [sp.chunks_x, sp.chunks_y],
           'kernel_launches_per_step': (sp.rp.launch_count() + sp.tp.launch_count() - l0) / steps,
           'bytes_sent_per_rank_per_step': sent, 'nvlink_GBps_per_rank': nvl,
           'nvlink_frac_of_900_nominal': nvl / NVLINK_NOMINAL, 'nvlink_frac_of_770_measured': nvl / NVLINK_MEASURED,
           'roofline_steps_per_s_at_900': NVLINK_NOMINAL * 1e9 / sent if sent else None,
           'hbm_algorithmic_GBps_per_rank': hbm, 'hbm_frac': hbm / hbm_peak()[0],
           'atom_number_first_last': [float(pops[0].sum()), float(pops[-1].sum())]}
    sp.close()
    return rec


# ----------------------------------------------------------------------------- the B200 arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    ctx = Ctx(rank, world, local_rank)
    dev = ctx.dev
    if world > 1:
        import datetime
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=300))
    cdtype = torch.complex128 if args.precision == 'c128' else torch.complex64
    mesh = args.mesh
    # each rank sweeps its own inter-component interaction (independent trajectories)
    ps = build_problem(mesh, g_ud=1.04 + 0.002 * rank, tag=f'rank{rank}')
    options = {}
    for name, val in (('col_tile', args.col_tile), ('row_mode', args.row_mode), ('stagger_ns', args.stagger_ns),
                      ('col_kernel', args.col_kernel), ('row_kernel', args.row_kernel)):
        if val is not None and val != 0:
            options[name] = val
    if args.graph is not None:
        options['graph'] = args.graph
    if args.no_prefetch:
        options['prefetch'] = 0
    for kv in args.opt or []:          # tuning runs: any sgpe_set_option selector
        k, v = kv.split('=')
        options[k] = int(v)

    pl = plan_for(ps, dev, args.precision, MODE, dense=args.dense, options=options)
    # (at least four warm-up steps: the steady-state step is captured into a CUDA graph from the fourth step of a call on,
    # and the capture belongs to the warm-up; the JSON line reports the number actually run)
    args.warmup = max(args.warmup, 4)
    pops = torch.zeros((1, args.steps + args.warmup, 2), dtype=torch.float64, device=dev)
    acct = pl.accounting()

    # ---- device-resident throughput ("value")
    pl.full_steps(args.warmup, pops, first=0)
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = pl.launch_count()
    ms_total = ctx.timed(lambda: pl.full_steps(args.steps, pops, first=args.warmup))
    sampler.stop_flag = True
    launches = pl.launch_count() - l0
    ms_step = ms_total / args.steps
    value = world * 1e3 / ms_step
    sampler.join(timeout=1.0)

    # ---- per-kernel durations for the roofline (event pairs around every pass, separate run)
    pl.profile_begin()
    pl.full_steps(args.steps, pops, first=args.warmup)
    prof = pl.profile_end()
    col_ms = prof['col_ms'] / max(1, prof['col_launches'])
    row_ms = prof['row_ms'] / max(1, prof['row_launches'])
    dom, dom_ms = ('col_pass', col_ms) if col_ms >= row_ms else ('row_pass', row_ms)
    per_launch = acct['algorithmic_bytes'] / 6.0
    peak, peak_src = hbm_peak()
    achieved = per_launch / (dom_ms * 1e-3) / 1e9

    # ---- energy tracking: E evaluated after every full step (the "energy tracking" of configs[2])
    n_e = max(1, min(args.steps, 50))
    eng = torch.zeros((1, n_e, 4), dtype=torch.float64, device=dev)
    # (first-use allocations, and - from four steps on - the capture of the replayed step, outside the timed region)
    pl.full_steps(min(5, n_e), pops, first=0, energy=eng, kl_term=2 * ps.kL_recoil)
    ms_energy_step = ctx.timed(lambda: pl.full_steps(n_e, pops, first=0, energy=eng, kl_term=2 * ps.kL_recoil)) / n_e

    # ---- end to end through the C ABI with host buffers: pinned state (+ the 1-D operator vectors, or the dense
    # grids under --dense) -> H2D -> K steps -> D2H state + populations.  The caller owns the pinned result buffers
    # (allocated once, as a user looping over runs would); three passes, the median is reported.
    psik_h = torch.as_tensor(np.array(ps.psik)).to(cdtype).reshape(1, 2, mesh, mesh).pin_memory()
    out_h = torch.empty_like(psik_h, pin_memory=True)
    pops_h = torch.zeros((1, args.steps, 2), dtype=torch.float64, pin_memory=True)
    pl2 = plan_for(ps, dev, args.precision, MODE, dense=args.dense, options=options)
    op_bytes = sum(t.numel() * t.element_size() for k, t in pl2.keep.items()
                   if k in ('kin', 'pot0', 'pot1', 'kin_x', 'kin_y', 'pot_x', 'pot_y'))
    from spinor_gpe_b200._separable import split_separable
    kin_np, pot_np = np.array(ps.kin_eng_spin), np.array(ps.pot_eng_spin)
    ksep = None if args.dense else split_separable(kin_np)
    psep = None if args.dense else split_separable(pot_np)

    def upload_operators(p):
        if ksep is not None:
            p.set_kinetic_separable(*ksep)
        else:
            p.set_kinetic(kin_np[0], kin_np[1])
        if psep is not None:
            p.set_potential_separable(*psep)
        else:
            p.set_potential(pot_np[0], pot_np[1], shared=True)

    pl2.run_host(psik_h, 2, want_pops=True)              # warm-up (allocations, first-use costs)
    e2e_runs = []
    for _ in range(3):
        ctx.barrier()
        t0 = time.perf_counter()
        upload_operators(pl2)
        pl2.run_host(psik_h, args.steps, want_pops=True, out=out_h, pops=pops_h)
        torch.cuda.synchronize(dev)
        e2e_runs.append(ctx.max_over_ranks((time.perf_counter() - t0) * 1e3))
    e2e_ms = float(np.median(e2e_runs))
    state_bytes = psik_h.numel() * psik_h.element_size()
    h2d = (state_bytes + op_bytes) / args.steps
    d2h = (state_bytes + pops_h.numel() * 8) / args.steps
    e2e_value = world * args.steps * 1e3 / e2e_ms
    atoms = float(pops_h[0, -1].sum())
    pl2.close()

    line = {
        'metric': 'full split-steps/s', 'value': value, 'unit': 'steps/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64' if args.precision == 'c128' else 'f32', 'data': 'synthetic',
        'config': workload_config(args), 'where': f'{world}x B200, one trajectory per GPU',
        'operators': pl.operators, 'kernel_options': options,
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                     'traffic': recorded_traffic(mesh), 'kernel': dom, 'kernel_ms': dom_ms,
                     'algorithmic_bytes_per_launch': per_launch, 'peak_source': peak_src,
                     'col_pass_ms': col_ms, 'row_pass_ms': row_ms,
                     'whole_step_frac': acct['algorithmic_bytes'] / (ms_step * 1e-3) / 1e9 / peak},
        'e2e': {'value': e2e_value, 'unit': 'steps/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'what': f'sgpe_run_host: pinned host operators+state -> {args.steps} full steps -> host state+pops; '
                        'median of 3 passes', 'ms_per_pass': [round(v, 3) for v in e2e_runs]},
        'gpu_launches': int(launches),
        'clocks': sampler.summary(),
        'energy_tracking': {'ms_per_step': ms_energy_step, 'value': world * 1e3 / ms_energy_step, 'unit': 'steps/s',
                            'frac': frac_of_roofline(1e3 / ms_energy_step, mesh * mesh, args.precision),
                            'what': 'full_step + eng_expect after every step (sgpe_full_steps_energy)',
                            'last_energy': [float(v) for v in eng[0, -1].cpu().numpy()]},
        'atom_number_check': atoms,
    }
    pl.close()
    del pl, pl2, psik_h, out_h
    torch.cuda.empty_cache()
    quick = args.quick

    # ---- the same mesh through the other kernel paths (device-resident, same timing method)
    def variants():
        out = {}
        for name, kw in (('dense_operators', dict(dense=True)), ('real_time', dict(mode='real')),
                         ('complex64', dict(precision='c64'))):
            p = plan_for(ps, dev, kw.get('precision', args.precision), kw.get('mode', MODE), dense=kw.get('dense', False),
                         options=options)
            rec, _ = steps_per_s(ctx, p, args.steps, min(args.warmup, 5))
            rec['value'] *= world
            rec['frac'] = frac_of_roofline(rec['value'] / world, mesh * mesh, kw.get('precision', args.precision))
            rec['operators'] = p.operators
            out[name] = rec
            p.close()
        return out

    # ---- configs[0] and configs[1]: launch-bound meshes
    def small_configs():
        out = {}
        ps0 = build_problem(256, tag=f'c0r{rank}')
        for graph in (0, 1):
            p = plan_for(ps0, dev, 'c128', 'imag', options=dict(options, graph=graph))
            rec, pp = steps_per_s(ctx, p, 400, 40, track_launches=True)
            rec['frac'] = frac_of_roofline(rec['value'], 256 * 256)
            rec['atom_number_check'] = float(pp[0, -1].sum())
            out['config0_256_ground_state' + ('_graph' if graph else '')] = rec
            p.close()
        ps1 = build_raman_problem(1024, tag=f'c1r{rank}')
        for graph in (0, 1):
            p = plan_for(ps1, dev, 'c128', 'real', options=dict(options, graph=graph))
            rec, pp = steps_per_s(ctx, p, 100, 10, track_launches=True)
            rec['frac'] = frac_of_roofline(rec['value'], 1024 * 1024)
            rec['atom_number_check'] = float(pp[0, -1].sum())
            out['config1_1024_raman_real_time' + ('_graph' if graph else '')] = rec
            p.close()
        return out

    # ---- the public drop-in: PSpinor.imaginary() as a user of the reference calls it (host arrays in and out,
    # PropResult with the final energy).  unwrap='herraez' is the reference's definition of eng_expect.
    def e2e_public():
        out = {}
        for unwrap in ('herraez', 'none'):
            q = build_problem(mesh, g_ud=1.04 + 0.002 * rank, tag=f'pub{rank}{unwrap}')
            q.imaginary(DT['imag'], 2, dev, unwrap=unwrap)                    # warm-up
            passes = []
            res = eng_final = None
            for rep in range(3):                                              # host-side work: median of three runs
                q = build_problem(mesh, g_ud=1.04 + 0.002 * rank, tag=f'pub{rank}{unwrap}b{rep}')
                # the previous result is dropped first: its page-locked staging blocks go back to torch's pool instead
                # of a fresh cudaHostAlloc of 268 MB (~0.3 s) landing in every other pass
                res = None
                ctx.barrier()
                t0 = time.perf_counter()
                res, _ = q.imaginary(DT['imag'], args.steps, dev, unwrap=unwrap)
                torch.cuda.synchronize(dev)
                passes.append(ctx.max_over_ranks((time.perf_counter() - t0) * 1e3))
                eng_final = [float(v) for v in res.eng_final]
            res = None
            ms = sorted(passes)[1]
            out[unwrap] = {'value': world * args.steps * 1e3 / ms, 'unit': 'steps/s', 'ms_total': ms,
                           'ms_per_pass': [round(v, 1) for v in passes], 'eng_final': eng_final}
        out['what'] = (f'PSpinor.imaginary(1/50, {args.steps}, "cuda"): host NumPy state and grids in, PropResult (psi, '
                       'psik, populations, final energy) out; herraez = phase-unwrapped energy as the reference defines it')
        return out

    if not quick:
        leg(line, 'variants', variants)
        leg(line, 'configs', small_configs)
        if MODE == 'imag':
            leg(line, 'e2e_public', e2e_public)
        leg(line, 'sweep', lambda: sweep_leg(ctx))
    parity_ok = True
    if world > 1 and not quick:
        leg(line, 'slab_parity', lambda: slab_parity_leg(ctx))
        leg(line, 'slab', lambda: slab_leg(ctx))
        if rank == 0:
            parity_ok = bool(line['slab_parity'].get('ok', False))

    if rank == 0:
        if world == 1 and not args.no_cpu and not quick:
            leg(line, 'library_bar', lambda: {
                'value': library_bar_steps_per_s(ps, dev), 'unit': 'steps/s',
                'what': 'the reference\'s op sequence (torch ATen + cuFFT, oracle restatement with device tensors) '
                        'on the same GPU, 10 full steps (benchmarks/benchmark_prop.py:58-97 method)'})
        if world == 1 and not args.no_cpu:
            sps, n, cores, med = oracle_steps_per_s(ps, 8, 1, budget_s=20.0)
            line['cpu_baseline'] = {'value': sps, 'unit': 'steps/s', 'cores': cores, 'kind': 'port',
                                    'sample': f'{n} full_step()+calc_pops of the same {mesh}^2 workload, median'}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not parity_ok:
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--mesh', type=int, default=2048)
    ap.add_argument('--precision', default='c128', choices=['c128', 'c64'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the CPU-baseline and library-bar legs (tuning runs)')
    ap.add_argument('--quick', action='store_true', help='headline numbers only (tuning runs)')
    ap.add_argument('--dense', action='store_true', help='force the general dense-operator path')
    ap.add_argument('--no-prefetch', action='store_true')
    ap.add_argument('--stagger-ns', type=int, default=0)
    ap.add_argument('--row-mode', type=int, default=0, choices=[0, 1])
    ap.add_argument('--col-tile', type=int, default=0, choices=[0, 2, 3, 8])
    ap.add_argument('--col-kernel', type=int, default=None, help='column-pass kernel selector (sgpe_set_option)')
    ap.add_argument('--row-kernel', type=int, default=None, help='row-pass kernel selector (sgpe_set_option)')
    ap.add_argument('--opt', action='append', help='name=value for sgpe_set_option (tuning runs; repeatable)')
    ap.add_argument('--graph', type=int, default=None, help='CUDA-graph replay of the steady-state step (0 / 1)')
    ap.add_argument('--mode', default='imag', choices=['imag', 'real'])
    args = ap.parse_args()
    global MODE
    MODE = args.mode
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if args.gpus != world and world == 1 and args.gpus > 1:
        print(json.dumps({'error': f'--gpus {args.gpus} needs torch.distributed.run with {args.gpus} ranks'}))
        sys.exit(2)
    run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
