"""Dev tool: first-contact check of the CUDA library on a GPU box (goldens + a 2048^2 timing)."""
import glob
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spinor_gpe_b200.plan import Plan   # noqa: E402
from spinor_gpe_b200 import _capi       # noqa: E402


def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


def plan_from_golden(z, pre, dtype=torch.complex128):
    g = lambda k: z[pre + 'in_' + k]   # noqa: E731
    psik = g('psik')
    ny, nx = psik.shape[-2:]
    pl = Plan(nx, ny, 1, dtype)
    dr = g('dr')
    pl.set_grid(dr[0], dr[1], float(g('dv_r')), float(g('dv_k')), float(g('atom_num')))
    pl.set_interactions(*g('g'))
    kin, pot = g('kin'), g('pot')
    pl.set_kinetic(kin[0], kin[1])
    pl.set_potential(pot[0], pot[1], shared=bool(np.array_equal(pot[0], pot[1])))
    if bool(g('is_coupling')):
        cpl = g('coupling')
        eiphi = None if bool(g('rot_coupling')) else np.exp(1j * 2 * float(g('kL')) * g('x'))
        if np.all(cpl == cpl.flat[0]):
            pl.set_coupling(_capi.SGPE_COUPLING_UNIFORM, omega=np.array([cpl.flat[0]]), eiphi=eiphi)
        else:
            pl.set_coupling(_capi.SGPE_COUPLING_DENSE, coupling=cpl, eiphi=eiphi)
    else:
        pl.set_coupling(_capi.SGPE_COUPLING_NONE)
    pl.set_time(str(z[pre + 'mode']), float(z[pre + 'dt']))
    pl.load(psik)
    return pl, psik


def goldens():
    worst = 0.0
    for f in sorted(glob.glob(os.path.join(ROOT, 'tests/golden/*.npz'))):
        if 'tensor_tools' in f:
            continue
        z = np.load(f)
        for r in range(int(z['n_runs'])):
            pre = f'r{r}_'
            n = int(z[pre + 'n_steps'])
            pl, psik = plan_from_golden(z, pre)
            dto, dti = pl.substeps()
            pl.single_step(dto)
            e1 = rel(pl.store()[0].cpu().numpy(), z[pre + 'psik_single_out'])
            pl.load(psik)
            pops = torch.zeros((1, n, 2), dtype=torch.float64, device='cuda')
            pl.full_steps(n, pops)
            e4 = rel(pl.store()[0].cpu().numpy(), z[pre + 'psik_final'])
            ep = float(np.abs(pops[0].cpu().numpy() - z[pre + 'pops_vals']).max() / np.abs(z[pre + 'pops_vals']).max())
            print(os.path.basename(f), r, 'single %.2e final(%d) %.2e pops %.2e' % (e1, n, e4, ep), flush=True)
            worst = max(worst, e1, e4, ep)
    print('WORST', worst)
    return worst


def timing(n=2048, mode='imag', steps=20, coupling=False, dtype=torch.complex128):
    nx = ny = n
    R = 8.0
    dx = 2 * R / n
    x = np.linspace(-R, R, n, endpoint=False)
    kx = np.linspace(-np.pi / dx, np.pi / dx, n, endpoint=False)
    X, Y = np.meshgrid(x, x)
    KX, KY = np.meshgrid(kx, kx)
    pot = (X ** 2 + Y ** 2) / 2
    kin = (KX ** 2 + KY ** 2) / 2
    psi = np.exp(-(X ** 2 + Y ** 2) / 4).astype(np.complex128)
    pl = Plan(nx, ny, 1, dtype)
    dvk = (np.pi / R) ** 2
    pl.set_grid(dx, dx, dx * dx, dvk, 100.0)
    pl.set_interactions(0.11, 0.11, 0.115)
    pl.set_kinetic(kin, kin)
    pl.set_potential(pot, pot, shared=True)
    if coupling:
        pl.set_coupling(_capi.SGPE_COUPLING_DENSE, coupling=np.zeros_like(pot))
    else:
        pl.set_coupling(_capi.SGPE_COUPLING_NONE)
    pl.set_time(mode, 1 / 50 if mode == 'imag' else 1 / 5000)
    st = torch.as_tensor(np.stack([psi, psi])[None])
    psik = pl.fft2d(st)
    psik = pl.normalise(psik, dvk)
    pl.load(psik)
    pops = torch.zeros((1, steps, 2), dtype=torch.float64, device='cuda')
    pl.full_steps(3, pops)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pl.full_steps(steps, pops)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    acc = pl.accounting()
    print(f'{n}^2 {mode} coupling={coupling} {dtype}: {ms:.3f} ms/full_step = {1e3 / ms:.1f} steps/s; '
          f'algorithmic {acc["algorithmic_bytes"] / ms * 1e-6:.0f} GB/s, designed traffic '
          f'{acc["actual_bytes"] / ms * 1e-6:.0f} GB/s; pops {pops[0, -1].tolist()}', flush=True)


if __name__ == '__main__':
    t0 = time.time()
    w = goldens()
    for n in (256, 1024, 2048, 4096):
        timing(n, 'imag')
    timing(2048, 'real')
    timing(2048, 'imag', coupling=True)
    timing(2048, 'real', coupling=True)
    timing(2048, 'imag', dtype=torch.complex64)
    print('elapsed', time.time() - t0)
    sys.exit(0 if w < 1e-10 else 1)
