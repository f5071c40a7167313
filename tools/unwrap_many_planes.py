"""Dev tool: phase unwrapping of MANY planes (the final energies of a parameter sweep): device-built spanning trees
(one plane after the other on the device) against the all-host merging (one plane per core)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from spinor_gpe_b200.plan import Plan

for n, planes in ((512, 128), (512, 16), (1024, 16), (256, 128)):
    pl = Plan(n, n)
    g = torch.Generator(device='cuda').manual_seed(1)
    f = torch.randn((planes, n, n), dtype=torch.float64, device='cuda', generator=g) \
        + 1j * torch.randn((planes, n, n), dtype=torch.float64, device='cuda', generator=g)
    y, x = torch.meshgrid(torch.arange(n, dtype=torch.float64, device='cuda'),
                          torch.arange(n, dtype=torch.float64, device='cuda'), indexing='ij')
    f[::2] = torch.polar(torch.ones_like(x), 0.011 * x + 0.007 * y + 6.0 * torch.sin(x / 97.0) * torch.cos(y / 131.0))
    ref = None
    for name, opts in (('default', {}), ('device tree + host anchor', {'unwrap_anchor': 0}), ('all-host merging', {'unwrap_merge': 1})):
        for k, v in {'unwrap_merge': 0, 'unwrap_anchor': -1}.items():
            pl.set_option(k, v)
        for k, v in opts.items():
            pl.set_option(k, v)
        out = pl.unwrap_phase(f)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = pl.unwrap_phase(f)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        ref = out if ref is None else ref
        print(f'{planes} planes of {n}^2, {name:28s}: {ms:8.1f} ms  ({"same" if torch.equal(out, ref) else "DIFFERENT"})', flush=True)
    pl.close()
