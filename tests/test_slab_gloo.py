"""Slab-decomposed propagation on two gloo ranks (CPU, emulated kernels): the all-to-all orchestration,
the transposed k-layout, the global norm all-reduce and the pack / unpack-transpose kernels must reproduce
the single-domain oracle."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, outdir, mode):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import spinor_oracle as orc
        from spinor_gpe_b200 import PSpinor
        from spinor_gpe_b200.slab import SlabPropagator
        from tests.emu_harness import emu_lib
        w0 = 2 * np.pi * 50
        ps = PSpinor(os.path.join(outdir, f'rank{rank}') + os.sep, overwrite=True, atom_num=1e4,
                     omeg={'x': w0, 'y': 1.5 * w0, 'z': 40 * w0}, g_sc={'uu': 1, 'dd': 0.98, 'ud': 1.02},
                     r_sizes=(16, 12), mesh_points=(128, 64))
        ps.coupling_setup(wavel=790.1e-9, kin_shift=True)
        ps.shift_momentum(scale=0.7, frac=(0.3, 0.7))
        ps.coupling_uniform(1.5 * ps.EL_recoil)
        ps.detuning_grad(-3.0)
        ps.rot_coupling = False
        dt, n = (1 / 50, 3) if mode == 'imag' else (1 / 2000, 3)
        prob = orc.Problem(ps.psik, ps.kin_eng_spin, ps.pot_eng_spin, ps.coupling, ps.space['dr'],
                           ps.space['dv_r'], ps.space['dv_k'], [ps.g_sc['uu'], ps.g_sc['dd'], ps.g_sc['ud']],
                           ps.atom_num, x=ps.space['x'], kL=ps.kL_recoil, is_coupling=True, rot_coupling=False)
        want = orc.OraclePropagator(prob, dt, mode).run(n)
        sp = SlabPropagator(ps, dt, time=mode, device='cpu', plan_kwargs={'_lib': emu_lib()})
        pops = torch.zeros((n, 2), dtype=torch.float64)
        sp.full_steps(n, pops)
        got = sp.gather_psik().numpy()
        err = np.linalg.norm(got - want['psik']) / np.linalg.norm(want['psik'])
        assert err < 1e-12, err
        np.testing.assert_allclose(pops.numpy(), want['pops_vals'], rtol=1e-12)
        assert sp.a2a_bytes > 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('mode', ['imag', 'real'])
def test_slab_two_gloo_ranks(mode):
    from tests.emu_harness import emu_lib
    emu_lib()
    mp.spawn(_worker, args=(2, _free_port(), tempfile.mkdtemp(prefix='sgpe_slab_'), mode), nprocs=2, join=True)
