"""Dev tool: phase times of one sgpe_unwrap_phase call (SGPE_UNWRAP_TIMING) on a noisy and a smooth 2048^2 plane."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['SGPE_UNWRAP_TIMING'] = '1'

import torch

from spinor_gpe_b200.plan import Plan

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
pl = Plan(n, n)
for name, value in (a.split('=') for a in sys.argv[2:]):
    pl.set_option(name, int(value))
g = torch.Generator(device='cuda').manual_seed(1)
f = torch.randn((n, n), dtype=torch.float64, device='cuda', generator=g) \
    + 1j * torch.randn((n, n), dtype=torch.float64, device='cuda', generator=g)
y, x = torch.meshgrid(torch.arange(n, dtype=torch.float64, device='cuda'),
                      torch.arange(n, dtype=torch.float64, device='cuda'), indexing='ij')
sm = torch.polar(torch.ones_like(x), 0.011 * x + 0.007 * y)
f2 = torch.stack([f, sm])
pl.unwrap_phase(f2)
torch.cuda.synchronize()
print('second call', file=sys.stderr, flush=True)
pl.unwrap_phase(f2)
