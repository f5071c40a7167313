"""Dev tool: per-CTA phase timeline of the row pass (globaltimer stamps)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from spinor_gpe_b200 import _lib
tl = os.path.join(ROOT, 'spinor_gpe_b200', 'libsgpe_timeline.so')     # make -C spinor_gpe_b200/csrc timeline
if os.path.exists(tl):
    _lib.LIB_PATH = tl
from spinor_gpe_b200 import _capi
from spinor_gpe_b200.plan import Plan, _dp
from spinor_gpe_b200._separable import split_separable
mesh = 2048
ps = bench.build_problem(mesh)
pl = Plan(mesh, mesh, 1)
pl.set_grid(ps.space['dr'][0], ps.space['dr'][1], ps.space['dv_r'], ps.space['dv_k'], ps.atom_num)
pl.set_interactions(ps.g_sc['uu'], ps.g_sc['dd'], ps.g_sc['ud'])
pl.set_kinetic(ps.kin_eng_spin[0], ps.kin_eng_spin[1])
pl.set_potential(ps.pot_eng_spin[0], ps.pot_eng_spin[1], shared=True)
pl.set_kinetic_separable(*split_separable(np.array(ps.kin_eng_spin)))
pl.set_potential_separable(*split_separable(np.array(ps.pot_eng_spin)))
pl.set_coupling(_capi.SGPE_COUPLING_NONE)
pl.set_time('imag', 1 / 50)
pl.load(np.array(ps.psik)[None])
kind = sys.argv[1] if len(sys.argv) > 1 else 'row'
for kv in filter(None, os.environ.get('SGPE_OPTS', '').split(',')):      # e.g. SGPE_OPTS=col_kernel=3
    k, v = kv.split('=')
    pl.set_option(k, int(v))
pl.full_steps(3)
ncta = mesh if kind == 'row' else 2 * mesh // 4
dbg = torch.zeros((ncta, 8), dtype=torch.int64, device='cuda')
pl.set_option('timeline_kind', 1 if kind == 'col' else 0)
dto, dti = pl.substeps()
pl.single_step(dto)                 # open a junction: the next column pass is the full one (FFT, factors, sums, iFFT)
pl.lib.sgpe_debug_timeline(pl.h, _dp(dbg))
pl.single_step(dti)
torch.cuda.synchronize()
pl.lib.sgpe_debug_timeline(pl.h, None)
d = dbg.cpu().numpy().astype(np.int64)
t0 = d[:, 0].min()
ph = d[:, :6] - t0
dur = np.diff(ph, axis=1)
names = (['issue loads', 'inverse FFT (incl. load wait)', 'point-wise', 'forward FFT', 'stores'] if kind == 'row' else
         ['issue loads + wait (persistent: wait for the staged tile)', 'forward FFT', 'K factors + publish sums', 'inverse FFT', 'stores'])
print('pass:', kind)
print('kernel span %.1f us, CTAs %d' % ((ph[:, 5].max()) / 1e3, len(d)))
for k, nme in enumerate(names):
    print('  %-32s mean %6.2f us  median %6.2f  p90 %6.2f' % (nme, dur[:, k].mean() / 1e3, np.median(dur[:, k]) / 1e3, np.percentile(dur[:, k], 90) / 1e3))
life = (ph[:, 5] - ph[:, 0]) / 1e3
print('  CTA lifetime mean %.2f us median %.2f' % (life.mean(), np.median(life)))
# per-SM overlap: fraction of time with 0/1/2 CTAs alive, and gaps between a CTA end and the next start
sm = d[:, 7]
gaps = []
for s in np.unique(sm):
    idx = np.where(sm == s)[0]
    ev = sorted([(ph[i, 0], ph[i, 5]) for i in idx])
    ends = sorted(e for _, e in ev)
    starts = sorted(st for st, _ in ev)
    # gap between k-th end and (k+2)-th start (2 slots per SM)
    per = 2 if kind == 'row' else 1
    for k in range(len(ends) - per):
        gaps.append(starts[k + per] - ends[k])
    if not gaps:
        gaps.append(0)
gaps = np.array(gaps) / 1e3
print('  slot turnaround (end of a CTA -> start of its successor on the SM): mean %.2f us median %.2f' % (gaps.mean(), np.median(gaps)))
print('  CTAs per SM: min %d max %d' % (np.bincount(sm).min(), np.bincount(sm).max()))
# phase alignment of co-resident CTAs: for a sample SM print the first 6 CTAs
s0 = np.unique(sm)[0]
idx = np.where(sm == s0)[0]
for i in sorted(idx, key=lambda i: ph[i, 0])[:8]:
    print('   SM%d cta %4d:' % (s0, i), ' '.join('%7.2f' % (x / 1e3) for x in ph[i]))
