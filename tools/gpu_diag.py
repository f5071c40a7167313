"""Dev tool: accuracy diagnostics on the GPU (per-size FFT error vs cuFFT, per-step propagation error)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spinor_gpe_b200.plan import Plan
from tests.test_gpu_parity import make_ps, problem_of, rel, SEEDED
from oracle import spinor_oracle as orc
from spinor_gpe_b200 import TensorPropagator

g = torch.Generator(device='cuda').manual_seed(1)
for n in (32, 64, 128, 256, 512, 1024, 2048, 4096):
    for (ny, nx) in ((n, 64), (64, n)):
        psi = torch.randn((1, 2, ny, nx), dtype=torch.float64, device='cuda', generator=g) + 1j * torch.randn((1, 2, ny, nx), dtype=torch.float64, device='cuda', generator=g)
        pl = Plan(nx, ny, 1)
        pl.set_grid(1.0, 1.0, 1.0, 1.0, 1.0)
        mine = pl.fft2d(psi)
        ref = torch.fft.fftshift(torch.fft.fftn(psi, dim=(-2, -1)) / (2 * np.pi), dim=(-2, -1))
        e = float((mine - ref).abs().max() / ref.abs().max())
        print(f'fft ny={ny} nx={nx} maxrel {e:.2e}', flush=True)

for mesh, mode, dt, n, cpl, rot, kshift in SEEDED[:4]:
    ps = make_ps(mesh, atom_num=1e4, r_sizes=(16, 16), g_sc={'uu': 1, 'dd': 0.98, 'ud': 1.02})
    if cpl != 'none':
        ps.coupling_setup(wavel=790.1e-9, kin_shift=kshift)
        if kshift:
            ps.shift_momentum(scale=0.7, frac=(0.3, 0.7))
        if cpl == 'uniform':
            ps.coupling_uniform(1.5 * ps.EL_recoil)
        elif cpl == 'dense':
            ps.coupling_grad(slope=0.3, offset=2.0, axis=1)
        ps.detuning_grad(-3.0)
    ps.rot_coupling = rot
    o = orc.OraclePropagator(problem_of(ps), dt, mode)
    prop = TensorPropagator(ps, dt, n, 'cuda', time=mode)
    errs = []
    for i in range(3):
        for which, dts in ((o.ops_out, prop.dt_out), (o.ops_in, prop.dt_in), (o.ops_out, prop.dt_out)):
            o.single_step(which)
            prop.single_step(dts)
            errs.append(rel(np.array([p.cpu().numpy() for p in prop.psik]), o.psik.numpy()))
    print(mesh, mode, cpl, ['%.1e' % e for e in errs], flush=True)
