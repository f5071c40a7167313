"""Dev tool: wall time of sgpe_unwrap_phase (device kernels + radix sort + host region merging) per mesh size,
device sort vs host sort, device-built spanning tree vs all-host merging, on smooth and noisy fields."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from spinor_gpe_b200.plan import Plan

for n in (512, 1024, 2048, 4096):
    pl = Plan(n, n)
    y, x = torch.meshgrid(torch.arange(n, dtype=torch.float64, device='cuda'),
                          torch.arange(n, dtype=torch.float64, device='cuda'), indexing='ij')
    smooth = torch.polar(torch.ones_like(x), 0.011 * x + 0.007 * y + 6.0 * torch.sin(x / 97.0) * torch.cos(y / 131.0))
    g = torch.Generator(device='cuda').manual_seed(1)
    noise = torch.randn((n, n), dtype=torch.float64, device='cuda', generator=g) \
        + 1j * torch.randn((n, n), dtype=torch.float64, device='cuda', generator=g)
    for name, f in (('smooth', smooth), ('noise', noise)):
        f2 = torch.stack([f, f.conj()])
        for sort, merge, anchor in ((0, 0, 1), (0, 0, 0), (0, 1, 0), (1, 1, 0)):
            if sort and n > 2048:
                continue
            pl.set_option('unwrap_sort', sort)
            pl.set_option('unwrap_merge', merge)
            pl.set_option('unwrap_anchor', anchor)
            pl.unwrap_phase(f2)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pl.unwrap_phase(f2)
            torch.cuda.synchronize()
            print(f'{n}^2 x 2 planes, {name:6s}, {"host" if sort else "device"} sort, '
                  f'{"all-host merging" if merge else ("device tree + device anchor bisection" if anchor else "device tree + host anchor pass")}: '
                  f'{(time.perf_counter() - t0) * 1e3:8.1f} ms', flush=True)
    pl.close()
