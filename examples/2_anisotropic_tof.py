"""Time of flight from an anisotropic trap (reference examples/2_anisotropic_tof.py): relax in imaginary time, switch
the trap off, expand in real time — the aspect ratio of the cloud inverts."""
import numpy as np

from _common import figures, options, report

args = options(mesh=512, steps_imag=1000, steps_real=1000)
from spinor_gpe_b200 import PSpinor      # noqa: E402

W = 2 * np.pi * 50
ps = PSpinor(args.data, overwrite=True, atom_num=1e4, omeg={'x': W, 'y': 4 * W, 'z': 40 * W},
             g_sc={'uu': 1, 'dd': 1, 'ud': 0.5}, phase_factor=1, pop_frac=(0.5, 0.5), r_sizes=(32, 32),
             mesh_points=(args.mesh, args.mesh))
ps.coupling_setup(wavel=790.1e-9)
ps.rand_seed = 99999
res0, _ = ps.imaginary(1 / 50, args.imag_steps, args.device)
report('trapped ground state', res0, ps)
figures(args, res0, rscale=ps.rad_tf, kscale=ps.kL_recoil, zoom=4)


def widths(dens):
    """rms widths (x, y) of the total density in units of a_x."""
    n = dens[0] + dens[1]
    x, y = ps.space['x_mesh'], ps.space['y_mesh']
    return float(np.sqrt((n * x ** 2).sum() / n.sum())), float(np.sqrt((n * y ** 2).sum() / n.sum()))


print('rms widths in the trap (x, y):', widths(res0.dens))
ps.pot_eng = np.zeros_like(ps.pot_eng)             # trap off
res1, _ = ps.real(1 / 500, args.real_steps, args.device, is_sampling=True, n_samples=min(50, args.real_steps))
report('after expansion', res1, ps)
print('rms widths after expansion (x, y):', widths(res1.dens))
figures(args, res1, rscale=ps.rad_tf, kscale=ps.kL_recoil, zoom=2)
