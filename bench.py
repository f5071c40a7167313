#!/usr/bin/env python
"""bench.py — full split-steps per second of the B200 propagator on BASELINE.json's headline config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mesh 2048]

Workload (BASELINE.json configs[2], the set-up of the reference's benchmarks/benchmark_prop.py:50-67):
imaginary-time ground state, 2048x2048 complex128, atom_num=1e2, omega=2pi*(50,50,2000), g=(1,1,1.04),
r_sizes=(8,8), coupling_setup(kin_shift=False) with Omega=0, dt=1/50, per-step renormalisation and per-step
populations; one "step" = one TensorPropagator.full_step() (3 split sub-steps).

N > 1 (launched by torch.distributed.run, one rank per GPU): every rank propagates its own independent
trajectory of the same mesh (a parameter sweep over g_ud), no data-path collective -> "scaling": "weak".

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for what each key means.
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
MODE = 'imag'                       # --mode real switches the whole run to real-time propagation
DT = {'imag': 1 / 50, 'real': 1 / 5000}


def build_problem(mesh, g_ud=1.04, tag='bench'):
    """PSpinor set-up of benchmark_prop.py (host NumPy, one-off)."""
    from spinor_gpe_b200 import PSpinor
    tmp = os.path.join(tempfile.mkdtemp(prefix='sgpe_bench_'), tag) + os.sep
    ps = PSpinor(tmp, overwrite=True, atom_num=1e2, omeg={'x': W0, 'y': W0, 'z': 40 * W0},
                 g_sc={'uu': 1, 'dd': 1, 'ud': g_ud}, pop_frac=(0.5, 0.5), r_sizes=(8, 8),
                 mesh_points=(mesh, mesh))
    ps.coupling_setup(wavel=790.1e-9, kin_shift=False)
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
            time.sleep(0.02)

    def summary(self):
        return {'sm_mhz': float(np.median(self.sm)) if self.sm else None, 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(self.sm)}


def oracle_steps_per_s(ps, max_steps, warmup, budget_s=40.0):
    """The CPU arm: the oracle's torch-CPU restatement of TensorPropagator.full_step on the host cores."""
    import torch
    from oracle import spinor_oracle as orc
    prob = orc.Problem(ps.psik, ps.kin_eng_spin, ps.pot_eng_spin, ps.coupling, ps.space['dr'],
                       ps.space['dv_r'], ps.space['dv_k'], [ps.g_sc['uu'], ps.g_sc['dd'], ps.g_sc['ud']],
                       ps.atom_num, x=ps.space['x'], kL=ps.kL_recoil, is_coupling=ps.is_coupling,
                       rot_coupling=ps.rot_coupling)
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
    single_step) on the SAME GPU: the oracle's restatement with its tensors moved to the device.  A baseline
    leg like cpu_baseline, reported under --library-bar only; never on the product path."""
    import torch
    from oracle import spinor_oracle as orc
    prob = orc.Problem(ps.psik, ps.kin_eng_spin, ps.pot_eng_spin, ps.coupling, ps.space['dr'],
                       ps.space['dv_r'], ps.space['dv_k'], [ps.g_sc['uu'], ps.g_sc['dd'], ps.g_sc['ud']],
                       ps.atom_num, x=ps.space['x'], kL=ps.kL_recoil, is_coupling=ps.is_coupling,
                       rot_coupling=ps.rot_coupling)
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


def run_reference(args, rank, world):
    if rank != 0:
        return
    ps = build_problem(args.mesh)
    sps, n, cores, med = oracle_steps_per_s(ps, args.steps, min(args.warmup, 2), budget_s=60.0)
    sample = f'{n} full_step() calls of the {args.mesh}x{args.mesh} workload after warm-up, median'
    line = {
        'impl': 'reference', 'metric': 'full split-steps/s', 'value': sps, 'unit': 'steps/s', 'n_gpus': 0,
        'steps': n, 'warmup': min(args.warmup, 2), 'ms_per_step': med * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'c128 (f64 arithmetic)', 'data': 'synthetic',
        'config': workload_config(args, 'cpu'),
        'cpu_baseline': {'value': sps, 'unit': 'steps/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': sps, 'unit': 'steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, where):
    kind = 'imaginary-time ground state' if MODE == 'imag' else 'real-time propagation (dt=1/5000)'
    return {'workload': f'{kind} {args.mesh}x{args.mesh} {args.precision}, benchmark_prop.py '
                        'parameters (atom_num=1e2, g=(1,1,1.04), r_sizes=(8,8), coupling_setup(kin_shift=False), '
                        'dt=1/50), per-step renormalisation + populations (BASELINE configs[2])',
            'mesh': [args.mesh, args.mesh], 'trajectories_per_gpu': 1, 'where': where,
            'l2': 'state (128 MiB at 2048^2) + operator grids exceed the 126 MB L2; no explicit flush'}


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from spinor_gpe_b200 import _capi
    from spinor_gpe_b200.plan import Plan

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if args.l2_fetch:
        import ctypes
        torch.cuda.init()
        rt = ctypes.CDLL('libcudart.so.12')
        rc = rt.cudaDeviceSetLimit(ctypes.c_int(0x05), ctypes.c_size_t(args.l2_fetch))   # cudaLimitMaxL2FetchGranularity
        got = ctypes.c_size_t()
        rt.cudaDeviceGetLimit(ctypes.byref(got), ctypes.c_int(0x05))
        print(f'l2 fetch granularity: rc={rc}, now {got.value} B', file=sys.stderr)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    cdtype = torch.complex128 if args.precision == 'c128' else torch.complex64
    mesh = args.mesh
    # each rank sweeps its own inter-component interaction (independent trajectories)
    ps = build_problem(mesh, g_ud=1.04 + 0.002 * rank, tag=f'rank{rank}')

    # ---- host (pinned) copies of everything the step needs
    psik_h = torch.as_tensor(np.array(ps.psik)).to(cdtype).reshape(1, 2, mesh, mesh).pin_memory()
    kin_h = torch.as_tensor(np.array(ps.kin_eng_spin)).pin_memory()
    pot_h = torch.as_tensor(np.array(ps.pot_eng_spin[0])).pin_memory()

    def make_plan():
        pl = Plan(mesh, mesh, 1, cdtype, dev)
        pl.set_grid(ps.space['dr'][0], ps.space['dr'][1], ps.space['dv_r'], ps.space['dv_k'], ps.atom_num)
        pl.set_interactions(ps.g_sc['uu'], ps.g_sc['dd'], ps.g_sc['ud'])
        pl.set_time(MODE, DT[MODE])
        pl.set_coupling(_capi.SGPE_COUPLING_NONE)        # Omega == 0: C is exactly the identity
        return pl

    from spinor_gpe_b200._separable import split_separable
    ksep = None if args.dense else split_separable(np.array(ps.kin_eng_spin))
    psep = None if args.dense else split_separable(np.array(ps.pot_eng_spin))

    def upload_operators(pl):
        kin_d = kin_h.to(dev, non_blocking=True)
        pot_d = pot_h.to(dev, non_blocking=True)
        pl.set_kinetic(kin_d[0], kin_d[1])
        pl.set_potential(pot_d, pot_d, shared=True)
        if ksep is not None:                         # separable fast path: 1-D factor tables
            pl.set_kinetic_separable(*ksep)
        if psep is not None:
            pl.set_potential_separable(*psep)
        if args.col_tile:
            pl.set_option('col_tile', args.col_tile)
        if args.row_mode:
            pl.set_option('row_mode', args.row_mode)
        if args.no_prefetch:
            pl.set_option('prefetch', 0)
        if args.stagger_ns:
            pl.set_option('stagger_ns', args.stagger_ns)

    pl = make_plan()
    upload_operators(pl)
    pl.load(psik_h.to(dev))
    pops = torch.zeros((1, args.steps + args.warmup, 2), dtype=torch.float64, device=dev)
    acct = pl.accounting()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- device-resident throughput ("value")
    pl.full_steps(args.warmup, pops, first=0)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    l0 = pl.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pl.full_steps(args.steps, pops, first=args.warmup)
    e1.record()
    barrier()
    sampler.stop_flag = True
    launches = pl.launch_count() - l0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
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

    # ---- energy tracking variant: E evaluated after every full step (extension of config 3)
    # (sgpe_full_steps_energy: the junction pass after every full step stores the boundary state on the side, its
    # inverse transform and the stencil pass run behind it; before that existed: one full_steps(1) + energy() pair
    # per step, timed as `separate_calls`)
    t_e0, t_e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_e = max(1, min(args.steps, 50))
    eng = torch.zeros((1, n_e, 4), dtype=torch.float64, device=dev)
    pl.full_steps(min(2, n_e), pops, first=0, energy=eng, kl_term=2 * ps.kL_recoil)    # first-use allocations
    barrier()
    t_e0.record()
    pl.full_steps(n_e, pops, first=0, energy=eng, kl_term=2 * ps.kL_recoil)
    t_e1.record()
    barrier()
    ms_energy_step = max_over_ranks(t_e0.elapsed_time(t_e1)) / n_e
    barrier()
    t_e0.record()
    for i in range(n_e):
        pl.full_steps(1, pops, first=i)
        pl.energy(None, kl_term=2 * ps.kL_recoil)
    t_e1.record()
    barrier()
    ms_energy_separate = max_over_ranks(t_e0.elapsed_time(t_e1)) / n_e

    # ---- end to end through the C ABI with host buffers (H2D of operators + state, steps, D2H)
    # The caller owns the pinned result buffers (allocated once, as a user looping over runs would); the timed region
    # is repeated three times and the median is reported (a cold first pass on a fresh box was seen 2x slower).
    pl2 = make_plan()
    upload_operators(pl2)
    out_h = torch.empty_like(psik_h, pin_memory=True)
    pops_h = torch.zeros((1, args.steps, 2), dtype=torch.float64, pin_memory=True)
    pl2.run_host(psik_h, 2, want_pops=True)              # warm-up (allocations, first-use costs)
    e2e_runs = []
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        upload_operators(pl2)
        pl2.run_host(psik_h, args.steps, want_pops=True, out=out_h, pops=pops_h)
        torch.cuda.synchronize(dev)
        e2e_runs.append(max_over_ranks((time.perf_counter() - t0) * 1e3))
    e2e_ms = float(np.median(e2e_runs))
    state_bytes = psik_h.numel() * psik_h.element_size()
    h2d = (state_bytes + kin_h.numel() * 8 + pot_h.numel() * 8) / args.steps
    d2h = (state_bytes + pops_h.numel() * 8) / args.steps
    e2e_value = world * args.steps * 1e3 / e2e_ms
    atoms = float(pops_h[0, -1].sum())

    if rank == 0:
        line = {
            'metric': 'full split-steps/s', 'value': value, 'unit': 'steps/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64' if args.precision == 'c128' else 'f32', 'data': 'synthetic',
            'config': dict(workload_config(args, f'{world}x B200, one trajectory per GPU'),
                           operators='separable factor tables (auto-detected)' if ksep is not None and psep is not None
                           else 'dense grids, factors evaluated per point', col_tile=args.col_tile),
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
            'energy_tracking': {'ms_per_step': ms_energy_step, 'value': world * 1e3 / ms_energy_step,
                                'what': 'full_step + eng_expect after every step (sgpe_full_steps_energy)',
                                'separate_calls': world * 1e3 / ms_energy_separate,
                                'last_energy': [float(v) for v in eng[0, -1].cpu().numpy()]},
            'atom_number_check': atoms,
        }
        # one evaluation of the energy the way the reference defines it (phase unwrapped): device kernels +
        # radix sort of the edges, region merging on the host, wall clock
        try:
            pl.energy(None, kl_term=2 * ps.kL_recoil, unwrap='herraez')
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            e_un = pl.energy(None, kl_term=2 * ps.kL_recoil, unwrap='herraez')[0].cpu().numpy()
            line['energy_unwrapped'] = {'ms': (time.perf_counter() - t0) * 1e3, 'energy': [float(v) for v in e_un],
                                        'what': 'one eng_expect with the reference\'s phase unwrapping'}
        except Exception as exc:                       # noqa: BLE001 - reported, never hides the headline numbers
            line['energy_unwrapped'] = {'error': str(exc)}
        if world == 1 and args.library_bar:
            line['library_bar'] = {'value': library_bar_steps_per_s(ps, dev), 'unit': 'steps/s',
                                   'what': 'the reference\'s op sequence (torch ATen + cuFFT, oracle restatement '
                                           'with device tensors) on the same GPU, 10 full steps'}
        if world == 1 and not args.no_cpu:
            sps, n, cores, med = oracle_steps_per_s(ps, 8, 1, budget_s=20.0)
            line['cpu_baseline'] = {'value': sps, 'unit': 'steps/s', 'cores': cores, 'kind': 'port',
                                    'sample': f'{n} full_step()+calc_pops of the same {mesh}^2 workload, median'}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--mesh', type=int, default=2048)
    ap.add_argument('--precision', default='c128', choices=['c128', 'c64'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the CPU-baseline leg (tuning runs)')
    ap.add_argument('--dense', action='store_true', help='force the general dense-operator path')
    ap.add_argument('--no-prefetch', action='store_true')
    ap.add_argument('--stagger-ns', type=int, default=0)
    ap.add_argument('--row-mode', type=int, default=0, choices=[0, 1])
    ap.add_argument('--col-tile', type=int, default=0, choices=[0, 2, 3, 8])
    ap.add_argument('--mode', default='imag', choices=['imag', 'real'])
    ap.add_argument('--l2-fetch', type=int, default=0, choices=[0, 32, 64, 128],
                    help='experiment: cudaLimitMaxL2FetchGranularity in bytes (0: leave the default)')
    ap.add_argument('--library-bar', action='store_true',
                    help='also time the reference op sequence (torch + cuFFT) on the same GPU')
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
