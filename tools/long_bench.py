"""Dev tool: full split-steps/s of ONE trajectory on ONE GPU for meshes beyond 4096 points per line
(slab.LongLinePlan: four-step lines, row-major k slab, no exchange partner).  Problem given by 1-D vectors
(slab.SeparableProblem), Thomas-Fermi state evaluated on the device."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from spinor_gpe_b200.slab import LongLinePlan, SeparableProblem

steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
chunk_list = [int(v) for v in sys.argv[3].split(',')] if len(sys.argv) > 3 else [None]
for n in [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else '8192,16384').split(',')]:
  for chunks in chunk_list:
    for mode, dt in (('imag', 1 / 50), ('real', 1 / 5000)):
        prob = SeparableProblem((n, n), (64, 64), atom_num=1e4, coupling=1.0, kin_shift=True, rot_coupling=False)
        pl = LongLinePlan(prob, dt, mode, 'cuda', chunks=chunks)
        pops = torch.zeros((1, steps + 2, 2), dtype=torch.float64, device='cuda')
        pl.full_steps(2, pops, first=0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pl.full_steps(steps, pops, first=2)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(json.dumps({'mesh': n, 'mode': mode, 'ms_per_step': round(ms, 3), 'steps_per_s': round(1e3 / ms, 2),
                          'algorithmic_GBps': round(768.0 * n * n / (ms * 1e-3) / 1e9, 1),
                          'atoms': float(pops[0, -1].sum()), 'n1': [pl.sp.n1x, pl.sp.n1y],
                          'chunks': [pl.sp.chunks_x, pl.sp.chunks_y]}), flush=True)
        pl.close()
        del pl, prob, pops
        torch.cuda.empty_cache()
