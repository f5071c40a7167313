"""Dev tool: time the building blocks separately at 2048^2 c128 (what does one transform / one memory pass cost?)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spinor_gpe_b200.plan import Plan

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
pl = Plan(n, n, 1)
pl.set_grid(1.0, 1.0, 1.0, 1.0, 100.0)
g = torch.Generator(device='cuda').manual_seed(0)
a = torch.randn((1, 2, n, n), dtype=torch.float64, device='cuda', generator=g) + 0j
b = torch.empty_like(a)

def timeit(f, reps=30):
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

import ctypes
from spinor_gpe_b200.plan import _dp
lib = pl.lib
res = {}
res['scale pass (read+write state, in place)'] = timeit(lambda: lib.sgpe_normalise(pl.h, _dp(a), _dp(a), 1.0, pl.stream)) 
res['sumsq only (read state)'] = timeit(lambda: pl.sumsq(a))
res['row FFT fwd in place (1 transform x2 comps)'] = timeit(lambda: lib.sgpe_fft1d(pl.h, _dp(a), _dp(a), 0, 0, pl.stream))
res['col FFT fwd in place (1 transform)'] = timeit(lambda: lib.sgpe_fft1d(pl.h, _dp(a), _dp(a), 1, 0, pl.stream))
res['fft2d out of place'] = timeit(lambda: lib.sgpe_fft2d(pl.h, _dp(a), _dp(b), 0, pl.stream))
res['torch copy a->b'] = timeit(lambda: b.copy_(a))
for k, v in res.items():
    print(f'{k:50s} {v*1e3:8.1f} us')
