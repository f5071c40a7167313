"""Dev tool: small propagation + energy + transforms for compute-sanitizer runs."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.gpu_diag2 import build
from spinor_gpe_b200 import TensorPropagator

for mesh, mode in (((512, 256), 'real'), ((128, 64), 'imag'), ((64, 2048), 'real')):
    ps = build(True, mesh)
    prop = TensorPropagator(ps, 1 / 2000, 2, 'cuda', time=mode)
    res = prop.prop_loop(2)
    print(mesh, mode, res.pops['vals'][-1], res.eng_final[0])
