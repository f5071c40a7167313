"""Dev tool: a few full steps of the headline workload, for ncu (keeps the profiled command short)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from spinor_gpe_b200 import _capi  # noqa: E402
from spinor_gpe_b200.plan import Plan  # noqa: E402

mesh = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = sys.argv[3] if len(sys.argv) > 3 else 'imag'
ps = bench.build_problem(mesh)
cdtype = torch.complex64 if os.environ.get('SGPE_PREC', 'c128') == 'c64' else torch.complex128
pl = Plan(mesh, mesh, 1, cdtype, torch.device('cuda', 0))
pl.set_grid(ps.space['dr'][0], ps.space['dr'][1], ps.space['dv_r'], ps.space['dv_k'], ps.atom_num)
pl.set_interactions(ps.g_sc['uu'], ps.g_sc['dd'], ps.g_sc['ud'])
pl.set_kinetic(ps.kin_eng_spin[0], ps.kin_eng_spin[1])
pl.set_potential(ps.pot_eng_spin[0], ps.pot_eng_spin[1], shared=True)
if os.environ.get('SGPE_DENSE', '0') != '1':
    from spinor_gpe_b200._separable import split_separable
    pl.set_kinetic_separable(*split_separable(np.array(ps.kin_eng_spin)))
    pl.set_potential_separable(*split_separable(np.array(ps.pot_eng_spin)))
if os.environ.get('SGPE_COL_TILE'):
    pl.set_option('col_tile', int(os.environ['SGPE_COL_TILE']))
for kv in filter(None, os.environ.get('SGPE_OPTS', '').split(',')):      # e.g. SGPE_OPTS=col_kernel=2,prefetch=0
    k, v = kv.split('=')
    pl.set_option(k, int(v))
pl.set_coupling(_capi.SGPE_COUPLING_NONE)
pl.set_time(mode, 1 / 50 if mode == 'imag' else 1 / 5000)
pl.load(np.array(ps.psik)[None].astype(np.complex64 if cdtype == torch.complex64 else np.complex128))
pops = torch.zeros((1, steps, 2), dtype=torch.float64, device='cuda')
pl.full_steps(steps, pops)
torch.cuda.synchronize()
print('pops', pops[0, -1].tolist())
