"""BASELINE config 4: 64 independent 512x512 trajectories (8 couplings x 8 detuning gradients of
examples/4_detuning_grad.py), imaginary time, sharded over the ranks (trajectory i -> rank i mod P), all
trajectories of a rank in ONE batched plan.  Prints aggregate trajectory-steps/s and the per-GPU HBM fraction.

    python tools/sweep_bench.py                                          # 1 GPU
    python -m torch.distributed.run --nproc-per-node P tools/sweep_bench.py
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from spinor_gpe_b200 import PSpinor, _capi  # noqa: E402
from spinor_gpe_b200._separable import split_separable  # noqa: E402
from spinor_gpe_b200.plan import Plan  # noqa: E402
from spinor_gpe_b200.sweep import detuning_coupling_grid, shard  # noqa: E402


def main():
    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    mesh, steps, warm = 512, int(sys.argv[1]) if len(sys.argv) > 1 else 50, 5
    w0 = 2 * np.pi * 50
    ps = PSpinor(os.path.join(tempfile.mkdtemp(prefix='sgpe_sweep_'), f'r{rank}') + os.sep, overwrite=True,
                 atom_num=1e4, omeg={'x': w0, 'y': w0, 'z': 40 * w0}, g_sc={'uu': 1, 'dd': 0.995, 'ud': 0.995},
                 pop_frac=(0.5, 0.5), r_sizes=(16, 16), mesh_points=(mesh, mesh))
    ps.coupling_setup(wavel=804e-9, kin_shift=True)
    ps.shift_momentum(scale=0.6, frac=(0.5, 0.5))
    trajs = detuning_coupling_grid(ps, np.linspace(0.5, 5, 8) * ps.EL_recoil, np.linspace(-12, 12, 8))
    mine = shard(len(trajs), rank, world)
    B = len(mine)
    pl = Plan(mesh, mesh, B, torch.complex128, dev)
    pl.set_grid(ps.space['dr'][0], ps.space['dr'][1], ps.space['dv_r'], ps.space['dv_k'], ps.atom_num)
    pl.set_interactions(ps.g_sc['uu'], ps.g_sc['dd'], ps.g_sc['ud'])
    kin = np.array(ps.kin_eng_spin)
    pl.set_kinetic(kin[0], kin[1])
    pl.set_kinetic_separable(*split_separable(kin))
    pots = np.stack([trajs[i].pot for i in mine])
    pl.set_potential(np.ascontiguousarray(pots[:, 0]), np.ascontiguousarray(pots[:, 1]), batched=True)
    seps = [split_separable(p) for p in pots]
    pl.set_potential_separable(np.stack([s[0] for s in seps]), np.stack([s[1] for s in seps]), batched=True)
    pl.set_coupling(_capi.SGPE_COUPLING_UNIFORM, omega=np.array([trajs[i].omega for i in mine]))
    pl.set_time('imag', 1 / 50)
    pl.load(np.stack([np.array(ps.psik)] * B))
    pops = torch.zeros((B, steps + warm, 2), dtype=torch.float64, device=dev)
    pl.full_steps(warm, pops)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pl.full_steps(steps, pops, first=warm)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    if rank == 0:
        total = len(trajs)
        tsps = total * 1e3 / ms
        per_gpu_bytes = 768.0 * mesh * mesh * B / (ms * 1e-3) / 1e9
        print(json.dumps({'sweep_bench': True, 'trajectories': total, 'mesh': mesh, 'ranks': world, 'per_rank_batch': B,
                          'ms_per_sweep_step': ms, 'trajectory_steps_per_s': tsps,
                          'hbm_algorithmic_GBps_per_gpu': per_gpu_bytes, 'hbm_frac_per_gpu': per_gpu_bytes / bench.hbm_peak()[0],
                          'pops_check': pops[0, -1].tolist()}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
