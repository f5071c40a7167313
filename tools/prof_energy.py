"""Dev tool: a few full steps with the energy of every step (for ncu on the energy-tracking chain)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ctx = bench.Ctx(0, 1, 0)
torch.cuda.set_device(0)
ps = bench.build_problem(2048)
pl = bench.plan_for(ps, ctx.dev)
n = 4
pops = torch.zeros((1, n, 2), dtype=torch.float64, device='cuda')
eng = torch.zeros((1, n, 4), dtype=torch.float64, device='cuda')
pl.full_steps(n, pops, energy=eng, kl_term=2 * ps.kL_recoil)
torch.cuda.synchronize()
print(eng[0, -1].tolist())
