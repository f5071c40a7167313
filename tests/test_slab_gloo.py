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


def _shared_exchange_buffers(world, mesh):
    """Exchange buffers of the fused ('p2p') variant for the CPU test: one k slab and one row slab per rank in shared
    memory, so that every process can store into every rank's buffers like the GPUs do through CUDA IPC."""
    n_local = 2 * mesh[0] * mesh[1] // world
    mk = lambda: [torch.zeros(n_local, dtype=torch.complex128).share_memory_() for _ in range(world)]   # noqa: E731
    return {'k': mk(), 'r': mk()}


def _worker(rank, world, port, outdir, mode, splits=(None, None), mesh=(128, 64), given=None, chunks=None):
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
                     r_sizes=(16, 12), mesh_points=mesh)
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
        sp = SlabPropagator(ps, dt, time=mode, device='cpu', plan_kwargs={'_lib': emu_lib()},
                            split_x=splits[0], split_y=splits[1], exchange='p2p' if given else 'nccl',
                            exchange_buffers=given, chunks=chunks)
        if chunks:
            assert max(sp.chunks_x, sp.chunks_y) == chunks
        pops = torch.zeros((n, 2), dtype=torch.float64)
        sp.full_steps(n, pops)
        got = sp.gather_psik().numpy()
        err = np.linalg.norm(got - want['psik']) / np.linalg.norm(want['psik'])
        assert err < 1e-12, err
        np.testing.assert_allclose(pops.numpy(), want['pops_vals'], rtol=1e-12)
        assert sp.a2a_bytes > 0
        if given and (mesh[0] * mesh[1] <= 128 * 64 or splits[1]):
            # distributed inverse transform to real space (ttools.ifft_2d of the state): this rank's rows, and the
            # k-space state is left untouched
            rows = sp.real_space_rows().numpy()
            nyl = mesh[1] // world
            want_rows = want['psi'][:, rank * nyl:(rank + 1) * nyl]
            err = np.linalg.norm(rows - want_rows) / np.linalg.norm(want_rows)
            assert err < 1e-12, err
            again = sp.gather_psik().numpy()
            assert np.linalg.norm(again - want['psik']) / np.linalg.norm(want['psik']) < 1e-12
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('mode', ['imag', 'real'])
def test_slab_two_gloo_ranks(mode):
    from tests.emu_harness import emu_lib
    emu_lib()
    mp.spawn(_worker, args=(2, _free_port(), tempfile.mkdtemp(prefix='sgpe_slab_'), mode), nprocs=2, join=True)


@pytest.mark.parametrize('splits,mesh', [((32, None), (1024, 64)), ((None, 32), (64, 1024))])
def test_slab_four_step_lines(splits, mesh):
    """The long-line machinery (four-step split, digit-transposed k order, strided pass with fused operators)
    forced on a small grid: 1024 = 32 x 32 along x or along y must still match the oracle."""
    from tests.emu_harness import emu_lib
    emu_lib()
    mp.spawn(_worker, args=(2, _free_port(), tempfile.mkdtemp(prefix='sgpe_slab_'), 'real', splits, mesh), nprocs=2,
             join=True)


@pytest.mark.parametrize('mode,splits,mesh', [('imag', (None, None), (128, 64)), ('real', (None, None), (128, 64)),
                                               ('real', (32, None), (1024, 64)), ('imag', (None, 32), (64, 1024))])
def test_slab_fused_exchange(mode, splits, mesh):
    """exchange='p2p': row-major k slab, column-wise k junction (kcol_pass / strided mid_pass) and the scatter stores
    of the last pass of each direction into the other rank's buffers, against the single-domain oracle."""
    from tests.emu_harness import emu_lib
    emu_lib()
    given = _shared_exchange_buffers(2, mesh)
    mp.spawn(_worker, args=(2, _free_port(), tempfile.mkdtemp(prefix='sgpe_slab_'), mode, splits, mesh, given),
             nprocs=2, join=True)


@pytest.mark.parametrize('mode,splits,mesh', [('real', (32, None), (1024, 64)), ('imag', (None, 32), (128, 1024))])
def test_slab_fused_exchange_chunked(mode, splits, mesh):
    """The chunked pipeline of the fused exchange (sgpe_slab_window): windows of the slab, per-chunk reduction slots,
    persistent scatter launches (3 CTAs walking each window in the emulation)."""
    from tests.emu_harness import emu_lib
    emu_lib()
    given = _shared_exchange_buffers(2, mesh)
    mp.spawn(_worker, args=(2, _free_port(), tempfile.mkdtemp(prefix='sgpe_slab_'), mode, splits, mesh, given, 2),
             nprocs=2, join=True)


def _worker_sep(rank, world, port, outdir, given=None):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from spinor_gpe_b200 import PSpinor
        from spinor_gpe_b200.slab import SeparableProblem, SlabPropagator
        from tests.emu_harness import emu_lib
        w0 = 2 * np.pi * 50
        kw = dict(atom_num=1e4, omeg={'x': w0, 'y': 1.5 * w0, 'z': 40 * w0}, g_sc={'uu': 1, 'dd': 0.98, 'ud': 1.02},
                  r_sizes=(16, 12), pop_frac=(0.5, 0.5))
        ps = PSpinor(os.path.join(outdir, f'rank{rank}') + os.sep, overwrite=True, mesh_points=(128, 64), **kw)
        ps.coupling_setup(wavel=790.1e-9, kin_shift=True)
        ps.coupling_uniform(1.5 * ps.EL_recoil)
        ps.detuning_grad(-3.0)
        sep = SeparableProblem((128, 64), coupling=1.5 * ps.EL_recoil, kin_shift=True, detuning_slope=-3.0, **kw)
        outs = []
        for prob in (ps, sep):
            sp = SlabPropagator(prob, 1 / 50, time='imag', device='cpu', plan_kwargs={'_lib': emu_lib()},
                                exchange='p2p' if given else 'nccl', exchange_buffers=given)
            sp.full_steps(2)
            outs.append(sp.gather_psik().numpy())
        err = np.linalg.norm(outs[1] - outs[0]) / np.linalg.norm(outs[0])
        assert err < 1e-11, err
    finally:
        dist.destroy_process_group()


def test_slab_separable_problem_matches_pspinor():
    """The host-light problem description (1-D vectors, Thomas-Fermi state generated per rank and transformed by
    the distributed FFT) reproduces the PSpinor-based set-up."""
    from tests.emu_harness import emu_lib
    emu_lib()
    mp.spawn(_worker_sep, args=(2, _free_port(), tempfile.mkdtemp(prefix='sgpe_slab_')), nprocs=2, join=True)


def test_slab_fused_exchange_real_space_setup():
    """The distributed forward transform of set_real_space through the scatter stores."""
    from tests.emu_harness import emu_lib
    emu_lib()
    given = _shared_exchange_buffers(2, (128, 64))
    mp.spawn(_worker_sep, args=(2, _free_port(), tempfile.mkdtemp(prefix='sgpe_slab_'), given), nprocs=2, join=True)
