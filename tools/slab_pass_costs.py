"""Dev tool (ONE GPU): time every local pass of the slab mode at the per-rank sizes of a MESH^2 / P run.
The process pretends to be rank 0 of P: torch.distributed is stubbed, the exchange buffers of all P "ranks" live on
this GPU, so a scatter store costs its HBM write but no NVLink.

    python tools/slab_pass_costs.py 16384 8 [p2p|nccl]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spinor_gpe_b200.slab import SeparableProblem, SlabPropagator  # noqa: E402

mesh = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
P = int(sys.argv[2]) if len(sys.argv) > 2 else 8
exchange = sys.argv[3] if len(sys.argv) > 3 else 'p2p'
dist.is_initialized = lambda: True
dist.get_rank = lambda group=None: 0
dist.get_world_size = lambda group=None: P
dist.all_reduce = lambda *a, **k: None
dev = torch.device('cuda', 0)
W0 = 2 * np.pi * 50
prob = SeparableProblem((mesh, mesh), r_sizes=(64, 64), atom_num=1e6, omeg={'x': W0, 'y': W0, 'z': 40 * W0},
                        g_sc={'uu': 1, 'dd': 1, 'ud': 1.04}, pop_frac=(0.5, 0.5), coupling=1.0, kin_shift=True,
                        rot_coupling=False)
n_local = 2 * mesh * mesh // P
given = None
if exchange == 'p2p':
    shared_k = torch.zeros(n_local, dtype=torch.complex128, device=dev)      # the other "ranks" share two dummies
    shared_r = torch.zeros(n_local, dtype=torch.complex128, device=dev)
    own_k = torch.zeros(n_local, dtype=torch.complex128, device=dev)
    own_r = torch.zeros(n_local, dtype=torch.complex128, device=dev)
    given = {'k': [own_k] + [shared_k] * (P - 1), 'r': [own_r] + [shared_r] * (P - 1)}
sp = SlabPropagator(prob, 1 / 5000, time='real', device=dev, exchange=exchange, exchange_buffers=given)
sp.sums[:] = torch.tensor([1.0, 0.5, 0.5, 0.0], device=dev) * (sp.atom_num / sp.dv_k)
print(f'mesh {mesh} P {P} exchange {exchange}: four-step splits x {sp.n1x} y {sp.n1y}; local slab '
      f'{n_local * 16 / 1e9:.2f} GB', flush=True)


def timeit(name, f, reps=5):
    f(); f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f'  {name:58s} {ms:7.3f} ms   {2 * n_local * 16 / ms * 1e-6:6.0f} GB/s (read + write of the slab)', flush=True)
    return ms


tp, rp, p2p = sp.tp, sp.rp, (exchange == 'p2p')
dt = sp.dt_in
tot = 0.0
if sp.n1y > 1:
    if p2p:
        tot += timeit('k: strided forward (mid_pass, inner)', lambda: tp.pass_mid(sp.tbuf, False, False, False, 0.0, True, True, None, 0.0, inner=sp.nxl))
        tot += timeit('k: contiguous fwd + K + inverse (kcol_pass)', lambda: tp.pass_kcols(sp.tbuf, True, False, 0.0, True, dt, True, sp.sums))
        tot += timeit('k: strided inverse + scatter (mid_pass, inner)', lambda: tp.pass_mid(sp.tbuf, True, True, False, 0.0, False, False, None, 0.0, inner=sp.nxl, scatter=True))
    else:
        tot += timeit('k: strided forward (mid_pass)', lambda: tp.pass_mid(sp.tbuf, False, False, False, 0.0, True, True, None, 0.0))
        tot += timeit('k: contiguous fwd + K + inverse (kline_pass)', lambda: tp.pass_klines(sp.tbuf, True, False, 0.0, True, dt, True, sp.sums))
        tot += timeit('k: strided inverse (mid_pass)', lambda: tp.pass_mid(sp.tbuf, True, True, False, 0.0, False, False, None, 0.0))
else:
    if p2p:
        tot += timeit('k: junction + scatter (kcol_pass)', lambda: tp.pass_kcols(sp.tbuf, True, False, 0.0, True, dt, True, sp.sums, scatter=True))
    else:
        tot += timeit('k: junction (kline_pass)', lambda: tp.pass_klines(sp.tbuf, True, False, 0.0, True, dt, True, sp.sums))
if sp.n1x > 1:
    tot += timeit('r: contiguous inverse (kline_pass)', lambda: rp.pass_klines(sp.rbuf, False, False, 0.0, False, 0.0, True, None))
    tot += timeit('r: strided inv + point-wise + strided fwd (mid_pass)', lambda: rp.pass_mid(sp.rbuf, True, True, True, dt, True, True, sp.sums, sp.points))
    tot += timeit('r: contiguous forward' + (' + scatter' if p2p else '') + ' (kline_pass)', lambda: rp.pass_klines(sp.rbuf, True, False, 0.0, False, 0.0, False, None, scatter=p2p))
else:
    tot += timeit('r: row pass' + (' + scatter' if p2p else ''), lambda: rp.pass_rows(sp.rbuf, dt, sp.sums, sp.points, scatter=p2p))
print(f'  sum of the local passes of one sub-step                    {tot:7.3f} ms')
