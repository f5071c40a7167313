"""Dev tool: replay the failing test path exactly, with intermediate checks."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_gpu_parity import make_ps, problem_of, rel
from tools.gpu_diag2 import build
from oracle import spinor_oracle as orc
from spinor_gpe_b200 import TensorPropagator

dt, n = 1 / 2000, 8
ps = build(True, (512, 256))
want = orc.OraclePropagator(problem_of(ps), dt, 'real').run(n)
g = lambda prop: np.array([p.cpu().numpy() for p in prop.psik])
# with pops
prop = TensorPropagator(ps, dt, n, 'cuda', time='real')
pops = torch.zeros((1, n, 2), dtype=torch.float64, device='cuda')
prop._plan.full_steps(n, pops)
print('with pops          %.1e' % rel(g(prop), want['psik']))
# energy then psik
prop = TensorPropagator(ps, dt, n, 'cuda', time='real')
prop._plan.full_steps(n, pops)
e = prop.eng_expect(None)
print('energy then psik   %.1e' % rel(g(prop), want['psik']), e, want['energy'])
# full prop_loop
prop = TensorPropagator(ps, dt, n, 'cuda', time='real')
res = prop.prop_loop(n)
print('prop_loop          %.1e' % rel(np.array(res.psik), want['psik']))
res2, prop2 = ps.real(dt, n, 'cuda')
print('ps.real            %.1e' % rel(np.array(res2.psik), want['psik']))
