"""Dev tool: small propagations + energy + transforms for compute-sanitizer runs — every kernel family once: the fused
passes, the persistent TMA-staged column passes (all variants), the graph replay, the generic-size passes, the streaming
energy kernel with and without the phase unwrapping."""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spinor_gpe_b200 import PSpinor, TensorPropagator  # noqa: E402


def build(mesh):
    w0 = 2 * np.pi * 50
    tmp = os.path.join(tempfile.mkdtemp(prefix='sgpe_san_'), 'run') + os.sep
    ps = PSpinor(tmp, overwrite=True, atom_num=1e4, omeg={'x': w0, 'y': w0, 'z': 40 * w0},
                 g_sc={'uu': 1, 'dd': 0.98, 'ud': 1.02}, pop_frac=(0.5, 0.5), r_sizes=(16, 16), mesh_points=mesh)
    ps.coupling_setup(wavel=790.1e-9, kin_shift=True)
    ps.coupling_uniform(1.5 * ps.EL_recoil)
    ps.rot_coupling = False
    return ps


cases = [((512, 256), 'real', 'c128', {}), ((128, 64), 'imag', 'c128', {'graph': 1}), ((64, 2048), 'real', 'c128', {}),
         ((64, 1024), 'imag', 'c128', {'col_kernel': 2}), ((64, 1024), 'real', 'c128', {'col_kernel': 3}),
         ((64, 1024), 'imag', 'c128', {'col_kernel': 4}), ((64, 1024), 'imag', 'c128', {'col_kernel': 5}),
         ((64, 1024), 'imag', 'c64', {'col_kernel': 6}), ((64, 1024), 'imag', 'c128', {'col_kernel': 7}),
         ((96, 80), 'real', 'c128', {}), ((30, 50), 'imag', 'c128', {})]
for mesh, mode, prec, opts in cases:
    ps = build(mesh)
    for sep in ((True, False) if opts.get('col_kernel', 2) == 2 else (True,)):
        prop = TensorPropagator(ps, 1 / 2000 if mode == 'real' else 1 / 50, 6, 'cuda', time=mode, precision=prec,
                                separable=sep, unwrap='none', track_energy=True)
        for k, v in opts.items():
            prop._plan.set_option(k, v)
        res = prop.prop_loop(6)
        print(mesh, mode, prec, opts, 'separable' if sep else 'dense', res.pops['vals'][-1], res.eng_final[0], flush=True)
from spinor_gpe_b200 import tensor_tools as tt  # noqa: E402
g = tt.grad([torch.as_tensor(np.asarray(p)).cuda() for p in build((96, 80)).psi], (0.1, 0.2))
print('gradient', float(g[0][0].abs().sum()), flush=True)
ps = build((128, 64))
prop = TensorPropagator(ps, 1 / 50, 2, 'cuda', time='imag')
print('unwrapped energy', prop.prop_loop(2).eng_final, flush=True)
torch.cuda.synchronize()
