"""Shared by the example scripts: command line, and figures only where matplotlib exists."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def options(mesh, steps_imag, steps_real=0):
    ap = argparse.ArgumentParser()
    ap.add_argument('--mesh', type=int, default=mesh, help='mesh points per axis (power of two)')
    ap.add_argument('--imag-steps', type=int, default=steps_imag)
    ap.add_argument('--real-steps', type=int, default=steps_real)
    ap.add_argument('--device', default='cuda')
    ap.add_argument('--data', default=None, help='data directory (default: a temporary one)')
    ap.add_argument('--plots', action='store_true', help='write the figures / movie (needs matplotlib, ffmpeg)')
    args = ap.parse_args()
    if args.data is None:
        import tempfile
        args.data = os.path.join(tempfile.mkdtemp(prefix='sgpe_example_'), 'trial') + os.sep
    return args


def figures(args, obj, **kw):
    """plot_spins / plot_total / plot_pops of a PSpinor or PropResult when --plots was given."""
    if not args.plots:
        return
    obj.plot_spins(**kw)
    for name in ('plot_total', 'plot_pops'):
        if hasattr(obj, name):
            getattr(obj, name)(**(kw if name == 'plot_total' else {}))


def report(tag, res, ps):
    print(f"{tag}: populations {res.pops['vals'][-1]}, atom number {res.pops['vals'][-1].sum():.6f}, "
          f"E_total {res.eng_final[0] * ps.space['dv_r']:.6f} hbar*omega_x (raw grid sum {res.eng_final[0]:.6e}), "
          f"phase separation {res.calc_separation():.4f}")
