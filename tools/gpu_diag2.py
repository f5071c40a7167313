"""Dev tool: isolate the (512,256) real-time discrepancy."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_gpu_parity import make_ps, problem_of, rel
from oracle import spinor_oracle as orc
from spinor_gpe_b200 import TensorPropagator

def build(noise, mesh=(512, 256)):
    ps = make_ps(mesh, atom_num=1e4, r_sizes=(16, 16), g_sc={'uu': 1, 'dd': 0.98, 'ud': 1.02})
    ps.coupling_setup(wavel=790.1e-9, kin_shift=True)
    ps.shift_momentum(scale=0.7, frac=(0.3, 0.7))
    ps.coupling_uniform(1.5 * ps.EL_recoil)
    ps.detuning_grad(-3.0)
    ps.rot_coupling = False
    if noise:
        rng = np.random.default_rng(99999)
        nz = [1 + 0.05 * (rng.standard_normal(p.shape) + 1j * rng.standard_normal(p.shape)) for p in ps.psik]
        ps.psik = [p * q for p, q in zip(ps.psik, nz)]
    return ps

if __name__ == "__main__":
 dt, n = 1 / 2000, 8
 for mesh in ((512, 256), (256, 256), (512, 512)):
   for noise in (False, True):
     ps = build(noise, mesh)
     o = orc.OraclePropagator(problem_of(ps), dt, 'real')
     wants = []
     for i in range(n):
         o.full_step(); wants.append(o.psik.numpy().copy())
     # (a) one call
     prop = TensorPropagator(ps, dt, n, 'cuda', time='real')
     prop._plan.full_steps(n)
     ea = rel(np.array([p.cpu().numpy() for p in prop.psik]), wants[-1])
     # (b) step by step, closing the junction each full step
     prop = TensorPropagator(ps, dt, n, 'cuda', time='real')
     eb = []
     for i in range(n):
         prop.full_step()
         eb.append(rel(np.array([p.cpu().numpy() for p in prop.psik]), wants[i]))
     # (c) repeat (a) to check determinism
     prop = TensorPropagator(ps, dt, n, 'cuda', time='real')
     prop._plan.full_steps(n)
     a2 = np.array([p.cpu().numpy() for p in prop.psik])
     prop = TensorPropagator(ps, dt, n, 'cuda', time='real')
     prop._plan.full_steps(n)
     a3 = np.array([p.cpu().numpy() for p in prop.psik])
     print(mesh, 'noise', noise, 'one-call %.1e' % ea, 'stepwise', ['%.0e' % e for e in eb], 'repeat-diff %.1e' % rel(a2, a3), flush=True)
