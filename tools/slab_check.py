"""Slab-mode check + timing under torch.distributed.run (one rank per GPU):
    python -m torch.distributed.run --nproc-per-node P tools/slab_check.py [mesh ...]
For every mesh: slab result after a few steps vs the single-GPU propagator (rank 0), then steps/s, all-to-all
bytes per step and the NVLink / HBM roofline fractions."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from spinor_gpe_b200 import TensorPropagator  # noqa: E402
from spinor_gpe_b200.slab import SlabPropagator  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    meshes = [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096]
    for mesh in meshes:
        ps = bench.build_problem(mesh, tag=f'slab{rank}')
        ps.coupling_uniform(0.5 * ps.EL_recoil)          # make the coupling operator non-trivial
        for mode, dt in (('imag', 1 / 50), ('real', 1 / 5000)):
            n = 3
            sp = SlabPropagator(ps, dt, time=mode, device=dev)
            pops = torch.zeros((n, 2), dtype=torch.float64, device=dev)
            sp.full_steps(n, pops)
            full = sp.gather_psik()
            if rank == 0:
                prop = TensorPropagator(ps, dt, n, dev, time=mode)
                pops1 = torch.zeros((1, n, 2), dtype=torch.float64, device=dev)
                prop._plan.full_steps(n, pops1)
                ref = torch.stack(prop.psik)
                err = float(torch.linalg.norm(full - ref) / torch.linalg.norm(ref))
                perr = float((pops - pops1[0]).abs().max() / pops1.abs().max())
                print(f'mesh {mesh} {mode}: slab({world} ranks) vs single GPU rel-L2 {err:.2e}, pops {perr:.2e}', flush=True)
                assert err < 1e-10 and perr < 1e-9
            del sp
        # timing (imaginary time)
        sp = SlabPropagator(ps, 1 / 50, time='imag', device=dev)
        steps = 20
        pops = torch.zeros((steps, 2), dtype=torch.float64, device=dev)
        sp.full_steps(3)
        dist.barrier(); torch.cuda.synchronize()
        sp.a2a_bytes = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sp.full_steps(steps, pops)
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / steps
        if rank == 0:
            sent = sp.a2a_bytes / steps                                  # bytes this rank sends per full step
            nvlink = sent / (ms * 1e-3) / 1e9
            hbm_alg = 768.0 * mesh * mesh / world / (ms * 1e-3) / 1e9
            print(json.dumps({'slab': True, 'mesh': mesh, 'ranks': world, 'ms_per_step': ms, 'steps_per_s': 1e3 / ms,
                              'a2a_bytes_sent_per_rank_per_step': sent, 'nvlink_GBps_per_rank': nvlink,
                              'nvlink_frac_of_770': nvlink / 770.0, 'hbm_algorithmic_GBps_per_rank': hbm_alg,
                              'hbm_frac': hbm_alg / bench.hbm_peak()[0]}), flush=True)
        del sp
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
