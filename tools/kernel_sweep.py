"""Dev tool: steps/s of the column-pass kernel variants over meshes / precisions / batches (one GPU)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ctx = bench.Ctx(0, 1, 0)
torch.cuda.set_device(0)
kernels = [int(k) for k in (sys.argv[1] if len(sys.argv) > 1 else '1,2,6').split(',')]
for mesh, steps in ((512, 200), (1024, 100), (2048, 40), (4096, 12)):
    ps = bench.build_problem(mesh)
    for prec in ('c128', 'c64'):
        for mode in ('imag', 'real'):
            row = {'mesh': mesh, 'precision': prec, 'mode': mode}
            for k in kernels:
                try:
                    pl = bench.plan_for(ps, ctx.dev, prec, mode, options={'col_kernel': k})
                    rec, _ = bench.steps_per_s(ctx, pl, steps, 5)
                    pl.profile_begin(); pl.full_steps(steps); prof = pl.profile_end()
                    row[f'k{k}'] = round(rec['value'], 1)
                    row[f'k{k}_col_us'] = round(1e3 * prof['col_ms'] / max(1, prof['col_launches']), 1)
                    pl.close()
                except Exception as exc:      # noqa: BLE001
                    row[f'k{k}'] = f'error: {exc}'
            print(json.dumps(row), flush=True)
# batched sweep (config 4 geometry): 8 and 64 trajectories of 512^2
from spinor_gpe_b200 import _capi  # noqa: E402
from spinor_gpe_b200._separable import split_separable  # noqa: E402
from spinor_gpe_b200.plan import Plan  # noqa: E402
ps = bench.build_problem(512)
for B in (8, 64):
    row = {'mesh': 512, 'batch': B}
    for k in kernels:
        pl = Plan(512, 512, B, torch.complex128, ctx.dev)
        pl.set_grid(ps.space['dr'][0], ps.space['dr'][1], ps.space['dv_r'], ps.space['dv_k'], ps.atom_num)
        pl.set_interactions(ps.g_sc['uu'], ps.g_sc['dd'], ps.g_sc['ud'])
        pl.set_kinetic_separable(*split_separable(np.array(ps.kin_eng_spin)))
        pl.set_potential_separable(*split_separable(np.array(ps.pot_eng_spin)))
        pl.set_coupling(_capi.SGPE_COUPLING_UNIFORM, omega=np.full(B, 0.3))
        pl.set_time('imag', 1 / 50)
        pl.set_option('col_kernel', k)
        pl.load(np.stack([np.array(ps.psik)] * B))
        pops = torch.zeros((B, 45, 2), dtype=torch.float64, device=ctx.dev)
        pl.full_steps(5, pops)
        ms = ctx.timed(lambda: pl.full_steps(40, pops, first=5)) / 40
        row[f'k{k}_traj_steps_per_s'] = round(B * 1e3 / ms, 1)
        pl.close()
    print(json.dumps(row), flush=True)
