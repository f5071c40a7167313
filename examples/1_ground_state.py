"""Ground state of a weakly interacting two-component condensate in a round harmonic trap (the set-up of the
reference's examples/1_ground_state.py): Thomas-Fermi guess -> imaginary time with sampling."""
import numpy as np

from _common import figures, options, report

args = options(mesh=256, steps_imag=100)
from spinor_gpe_b200 import PSpinor      # noqa: E402

W = 2 * np.pi * 50
ps = PSpinor(args.data, overwrite=True, atom_num=1e2, omeg={'x': W, 'y': W, 'z': 40 * W},
             g_sc={'uu': 1, 'dd': 1, 'ud': 1.04}, pop_frac=(0.5, 0.5), r_sizes=(8, 8),
             mesh_points=(args.mesh, args.mesh))
ps.coupling_setup(wavel=790.1e-9, kin_shift=False)
figures(args, ps, rscale=ps.rad_tf, kscale=ps.kL_recoil, zoom=4)
ps.rand_seed = 99999
res, prop = ps.imaginary(1 / 50, args.imag_steps, args.device, is_sampling=True, n_samples=min(50, args.imag_steps))
report('ground state', res, ps)
print('kinetic energy per component (spectral):', prop.kin_expect_spectral())
figures(args, res, rscale=ps.rad_tf, kscale=ps.kL_recoil, zoom=4)
if args.plots:
    res.make_movie(rscale=ps.rad_tf, kscale=ps.kL_recoil, zoom=4, norm_type='half')
