"""GPU tests of the slab-decomposed mode (``-m gpu``, need >= 2 devices; skipped otherwise): one process per GPU
over NCCL, the slab result against the single-GPU propagator and the CPU oracle on the same problem, for both
exchange variants (fused scatter stores through CUDA-IPC peer memory / pack + NCCL all-to-all + unpack) and with a
forced four-step split of the lines."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, exchange, splits, mesh, mode):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        from oracle import spinor_oracle as orc
        from spinor_gpe_b200 import PSpinor, TensorPropagator
        from spinor_gpe_b200.slab import SlabPropagator
        w0 = 2 * np.pi * 50
        ps = PSpinor(os.path.join(tempfile.mkdtemp(prefix='sgpe_gslab_'), f'r{rank}') + os.sep, overwrite=True,
                     atom_num=1e3, omeg={'x': w0, 'y': w0, 'z': 40 * w0}, g_sc={'uu': 1, 'dd': 1, 'ud': 1.04},
                     pop_frac=(0.5, 0.5), r_sizes=(8, 8), mesh_points=mesh)
        ps.coupling_setup(wavel=790.1e-9, kin_shift=True)
        ps.coupling_uniform(0.5 * ps.EL_recoil)
        ps.rot_coupling = False
        dt, n = (1 / 50, 3) if mode == 'imag' else (1 / 2000, 3)
        sp = SlabPropagator(ps, dt, time=mode, device=dev, exchange=exchange, split_x=splits[0], split_y=splits[1])
        assert sp.exchange == exchange
        pops = torch.zeros((n, 2), dtype=torch.float64, device=dev)
        sp.full_steps(n, pops)
        got = sp.gather_psik()
        if rank == 0:
            prop = TensorPropagator(ps, dt, n, dev, time=mode)
            pops1 = torch.zeros((1, n, 2), dtype=torch.float64, device=dev)
            prop._plan.full_steps(n, pops1)
            one = torch.stack(prop.psik)
            err_gpu = float(torch.linalg.norm(got - one) / torch.linalg.norm(one))
            prob = orc.Problem(ps.psik, ps.kin_eng_spin, ps.pot_eng_spin, ps.coupling, ps.space['dr'],
                               ps.space['dv_r'], ps.space['dv_k'], [ps.g_sc['uu'], ps.g_sc['dd'], ps.g_sc['ud']],
                               ps.atom_num, x=ps.space['x'], kL=ps.kL_recoil, is_coupling=True, rot_coupling=False)
            want = orc.OraclePropagator(prob, dt, mode).run(n)
            err = np.linalg.norm(got.cpu().numpy() - want['psik']) / np.linalg.norm(want['psik'])
            perr = np.abs(pops.cpu().numpy() - want['pops_vals']).max() / np.abs(want['pops_vals']).max()
            assert err_gpu < 1e-10 and err < 1e-10 and perr < 1e-9, (err_gpu, err, perr)
        dist.barrier()
        sp.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('exchange,splits,mesh,mode', [
    ('p2p', (None, None), (256, 128), 'imag'),
    ('p2p', (None, None), (256, 128), 'real'),
    ('p2p', (32, 32), (1024, 1024), 'real'),
    ('nccl', (None, None), (256, 128), 'imag'),
    ('nccl', (32, None), (1024, 256), 'real'),
])
def test_slab_matches_single_gpu_and_oracle(exchange, splits, mesh, mode):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs')
    mp.spawn(_worker, args=(2, _free_port(), exchange, splits, mesh, mode), nprocs=2, join=True)
