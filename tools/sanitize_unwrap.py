"""Dev tool: the device-side region merging of the phase unwrapping (Boruvka rounds, anchor bisection down to single
pixels, per-step energy in the reference's definition) on small planes, for compute-sanitizer."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from spinor_gpe_b200.plan import Plan

for ny, nx in ((128, 256), (96, 80)):
    pl = Plan(nx, ny)
    rng = np.random.default_rng(3)
    y, x = np.mgrid[0:ny, 0:nx].astype(np.float64)
    vort = (x - 0.31 * nx + 1j * (y - 0.37 * ny)) * (x - 0.62 * nx - 1j * (y - 0.71 * ny)) * np.exp(0.2j * x)
    noise = rng.normal(size=(ny, nx)) + 1j * rng.normal(size=(ny, nx))
    planes = torch.as_tensor(np.stack([vort + 0.5 * noise, noise, np.exp(0.3j * x), np.ones((ny, nx), dtype=complex)])).cuda()
    ref = None
    for opts in ({'unwrap_merge': 1}, {'unwrap_anchor': 0}, {'unwrap_anchor': 1, 'unwrap_tail': 0},
                 {'unwrap_anchor': 1, 'unwrap_tail': 500}):
        for k in ('unwrap_merge', 'unwrap_anchor', 'unwrap_tail'):
            pl.set_option(k, {'unwrap_merge': 0, 'unwrap_anchor': -1, 'unwrap_tail': 16384}[k])
        for k, v in opts.items():
            pl.set_option(k, v)
        out = pl.unwrap_phase(planes)
        ref = out if ref is None else ref
        print((ny, nx), opts, 'same field' if torch.equal(out, ref) else 'DIFFERENT', flush=True)
    pl.close()
torch.cuda.synchronize()
