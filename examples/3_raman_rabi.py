"""Raman-driven Rabi flopping (reference examples/3_raman_rabi.py): all atoms in one spin state, momentum kick of one
recoil, short imaginary-time relaxation without coupling, then real time with a uniform Raman coupling in the
laboratory frame — the populations oscillate between the two components."""
import numpy as np

from _common import figures, options, report

args = options(mesh=256, steps_imag=1000, steps_real=2000)
from spinor_gpe_b200 import PSpinor      # noqa: E402

W = 2 * np.pi * 50
ps = PSpinor(args.data, overwrite=True, atom_num=1e4, omeg={'x': W, 'y': W, 'z': 40 * W},
             g_sc={'uu': 1, 'dd': 1, 'ud': 0.0}, pop_frac=(1.0, 0.0), r_sizes=(16, 16),
             mesh_points=(args.mesh, args.mesh))
ps.coupling_setup(wavel=790.1e-9, kin_shift=True)
ps.shift_momentum(scale=1.0, frac=(0, 1.0))
ps.rot_coupling = False
ps.rand_seed = 99999
res0, _ = ps.imaginary(1 / 50, args.imag_steps, args.device, is_sampling=True, n_samples=min(50, args.imag_steps))
report('relaxed', res0, ps)
ps.coupling_uniform(1.0 * ps.EL_recoil)
res1, _ = ps.real(1 / 5000, args.real_steps, args.device, is_sampling=True, n_samples=min(100, args.real_steps))
report('driven', res1, ps)
up = res1.pops['vals'][:, 0]
print(f'spin-up population: min {up.min():.1f}, max {up.max():.1f} of {ps.atom_num:.0f} over '
      f"{res1.pops['times'][-1] * ps.time_scale * 1e3:.3f} ms")
figures(args, res1, rscale=ps.rad_tf, kscale=ps.kL_recoil, zoom=1)
