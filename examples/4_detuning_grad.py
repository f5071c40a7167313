"""Ground state in a Raman detuning gradient (reference examples/4_detuning_grad.py): uniform coupling plus a linear
detuning along y separates the components vertically; a row of vortices forms where they meet.  With --sweep the
same set-up runs as a batched parameter sweep over coupling strength and gradient (BASELINE config 4)."""
import sys

import numpy as np

from _common import figures, options, report

sweep = '--sweep' in sys.argv
if sweep:
    sys.argv.remove('--sweep')
args = options(mesh=256, steps_imag=1000)
from spinor_gpe_b200 import PSpinor      # noqa: E402

W = 2 * np.pi * 50
ps = PSpinor(args.data, overwrite=True, atom_num=1e4, omeg={'x': W, 'y': W, 'z': 40 * W},
             g_sc={'uu': 1, 'dd': 0.995, 'ud': 0.995}, pop_frac=(0.5, 0.5), r_sizes=(16, 16),
             mesh_points=(args.mesh, args.mesh))
ps.coupling_setup(wavel=804e-9, kin_shift=True)
ps.shift_momentum(scale=0.6, frac=(0.5, 0.5))
if not sweep:
    ps.coupling_uniform(5 * ps.EL_recoil)
    ps.detuning_grad(-12)
    res, _ = ps.imaginary(1 / 50, args.imag_steps, args.device, is_sampling=True, n_samples=min(50, args.imag_steps))
    report('detuning gradient', res, ps)
    figures(args, res, rscale=ps.rad_tf, kscale=ps.kL_recoil, zoom=2)
else:
    from spinor_gpe_b200.sweep import detuning_coupling_grid, run_sweep
    couplings = np.linspace(0.5, 5, 4) * ps.EL_recoil
    slopes = np.linspace(-12, 12, 4)
    out = run_sweep(ps, detuning_coupling_grid(ps, couplings, slopes), 1 / 50, args.imag_steps, time='imag',
                    device=args.device, batch=8)
    for i, (c, s) in enumerate((c, s) for c in couplings for s in slopes):
        print(f'Omega = {c / ps.EL_recoil:4.2f} E_L, slope = {s:6.1f}: populations {out["pops"][i, -1]}')
