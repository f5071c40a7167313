"""N > 1 host logic on CPU: two gloo ranks shard a parameter sweep (round-robin), run their trajectories
(here on the emulated kernels) and gather; the result must equal a loop of single oracle runs.  Also the
pure sharding arithmetic."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spinor_gpe_b200.sweep import shard


def test_shard_partition():
    for n in (1, 7, 64):
        for world in (1, 2, 3, 8):
            parts = [shard(n, r, world) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, outdir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import spinor_oracle as orc
        from spinor_gpe_b200 import PSpinor
        from spinor_gpe_b200.sweep import detuning_coupling_grid, run_sweep
        from tests.emu_harness import EmuTorchPlan
        w0 = 2 * np.pi * 50
        ps = PSpinor(os.path.join(outdir, f'rank{rank}') + os.sep, overwrite=True, atom_num=1e4,
                     omeg={'x': w0, 'y': w0, 'z': 40 * w0}, g_sc={'uu': 1, 'dd': 0.995, 'ud': 0.995},
                     r_sizes=(16, 16), mesh_points=(32, 32))
        ps.coupling_setup(wavel=804e-9, kin_shift=True)
        ps.shift_momentum(scale=0.6, frac=(0.5, 0.5))
        trajs = detuning_coupling_grid(ps, [0.5 * ps.EL_recoil, 5 * ps.EL_recoil], [-12.0, 0.0, 12.0])
        out = run_sweep(ps, trajs, 1 / 50, 3, time='imag', device='cpu', batch=2, keep_states=True,
                        plan_factory=EmuTorchPlan)
        assert out['pops'].shape == (6, 3, 2) and out['energy'].shape == (6, 4)
        np.testing.assert_array_equal(out['owner'], np.arange(6) % world)
        if rank == 0:
            for i, tr in enumerate(trajs):
                prob = orc.Problem(ps.psik, ps.kin_eng_spin, tr.pot, np.full_like(ps.pot_eng, tr.omega),
                                   ps.space['dr'], ps.space['dv_r'], ps.space['dv_k'],
                                   [ps.g_sc['uu'], ps.g_sc['dd'], ps.g_sc['ud']], ps.atom_num, x=ps.space['x'],
                                   kL=ps.kL_recoil, is_coupling=True, rot_coupling=True)
                want = orc.OraclePropagator(prob, 1 / 50, 'imag').run(3)
                err = np.linalg.norm(out['psik'][i] - want['psik']) / np.linalg.norm(want['psik'])
                assert err < 1e-12, (i, err)
                np.testing.assert_allclose(out['pops'][i], want['pops_vals'], rtol=1e-12)
    finally:
        dist.destroy_process_group()


def test_sweep_two_gloo_ranks():
    outdir = tempfile.mkdtemp(prefix='sgpe_gloo_')
    from tests.emu_harness import emu_lib
    emu_lib()                                   # build once, before forking
    mp.spawn(_worker, args=(2, _free_port(), outdir), nprocs=2, join=True)
