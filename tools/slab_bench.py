"""Slab-mode validation + benchmark under torch.distributed.run (one rank per GPU).

    python -m torch.distributed.run --nproc-per-node P tools/slab_bench.py check      # parity of the four-step path
    python -m torch.distributed.run --nproc-per-node P tools/slab_bench.py bench MESH [steps]

check: 2048^2 with forced 32x64 / 64x32 splits and the natural path vs the single-GPU propagator (rank 0).
bench: BASELINE config 5 style — real-time propagation with uniform Raman coupling on a MESH^2 grid
       (16384 on 8 GPUs), set up from 1-D vectors only; prints steps/s, all-to-all bytes and the NVLink / HBM
       roofline fractions (algorithmic 768 B/pt/step; NVLink 770 GB/s per direction measured on this pool)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from spinor_gpe_b200.slab import SeparableProblem, SlabPropagator  # noqa: E402

W0 = 2 * np.pi * 50
CHUNKS = None              # --chunks=N: windows of the chunked exchange pipeline (default: SlabPropagator's choice)
CTAS = None                # --ctas=N: persistent CTAs of a scatter launch
EXCHANGE = 'p2p'           # --exchange=nccl: pack / NCCL all-to-all / unpack instead of the fused scatter stores


def check(rank, world, dev):
    from spinor_gpe_b200 import TensorPropagator
    mesh = 2048
    ps = bench.build_problem(mesh, tag=f'slab{rank}')
    ps.coupling_uniform(0.5 * ps.EL_recoil)
    ps.rot_coupling = False
    ref = None
    for mode, dt in (('real', 1 / 5000), ('imag', 1 / 50)):
        n = 2
        if rank == 0:
            prop = TensorPropagator(ps, dt, n, dev, time=mode)
            prop._plan.full_steps(n)
            ref = torch.stack(prop.psik)
        for splits in ((None, None), (32, None), (None, 64), (64, 32)):
            sp = SlabPropagator(ps, dt, time=mode, device=dev, split_x=splits[0], split_y=splits[1], exchange=EXCHANGE, chunks=CHUNKS, scatter_ctas=CTAS)
            pops = torch.zeros((n, 2), dtype=torch.float64, device=dev)
            sp.full_steps(n, pops)
            full = sp.gather_psik()
            if rank == 0:
                err = float(torch.linalg.norm(full - ref) / torch.linalg.norm(ref))
                print(f'check {mesh}^2 {mode} splits={splits} ranks={world} exchange={EXCHANGE}: rel-L2 vs single GPU {err:.2e} '
                      f'atoms {float(pops[-1].sum()):.10g}', flush=True)
                assert err < 1e-10
            sp.close()
            del sp


def run_bench(rank, world, dev, mesh, steps):
    g_sc = {'uu': 1, 'dd': 1, 'ud': 1.04}
    prob = SeparableProblem((mesh, mesh), r_sizes=(64, 64), atom_num=1e6, omeg={'x': W0, 'y': W0, 'z': 40 * W0},
                            g_sc=g_sc, pop_frac=(0.5, 0.5), coupling=1.0, kin_shift=True, rot_coupling=False)
    sp = SlabPropagator(prob, 1 / 5000, time='real', device=dev, exchange=EXCHANGE, chunks=CHUNKS, scatter_ctas=CTAS)
    pops = torch.zeros((steps, 2), dtype=torch.float64, device=dev)
    sp.full_steps(2)
    dist.barrier(); torch.cuda.synchronize()
    sp.a2a_bytes = 0
    l0 = sp.rp.launch_count() + sp.tp.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sp.full_steps(steps, pops)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    if rank == 0:
        sent = sp.a2a_bytes / steps
        nvl = sent / (ms * 1e-3) / 1e9
        hbm = 768.0 * mesh * mesh / world / (ms * 1e-3) / 1e9
        print(json.dumps({
            'slab_bench': True, 'mesh': mesh, 'ranks': world, 'mode': 'real', 'dtype': 'c128', 'exchange': EXCHANGE,
            'four_step': [sp.n1x, sp.n1y], 'chunks': [sp.chunks_x, sp.chunks_y], 'scatter_ctas': sp.scatter_ctas, 'steps': steps, 'ms_per_step': ms, 'steps_per_s': 1e3 / ms,
            'kernel_launches_per_step': (sp.rp.launch_count() + sp.tp.launch_count() - l0) / steps,
            'a2a_bytes_sent_per_rank_per_step': sent, 'nvlink_GBps_per_rank': nvl, 'nvlink_frac_of_770': nvl / 770.0,
            'hbm_algorithmic_GBps_per_rank': hbm, 'hbm_frac': hbm / bench.hbm_peak()[0],
            'atoms_first_last': [float(pops[0].sum()), float(pops[-1].sum())],
            'pops_last': pops[-1].tolist()}), flush=True)


def breakdown(rank, world, dev, mesh):
    """Time the phases of one sub-step with CUDA events (rank 0 prints; max over ranks not taken: indicative)."""
    prob = SeparableProblem((mesh, mesh), r_sizes=(64, 64), atom_num=1e6, omeg={'x': W0, 'y': W0, 'z': 40 * W0},
                            g_sc={'uu': 1, 'dd': 1, 'ud': 1.04}, pop_frac=(0.5, 0.5), coupling=1.0, kin_shift=True,
                            rot_coupling=False)
    sp = SlabPropagator(prob, 1 / 5000, time='real', device=dev, exchange=EXCHANGE, chunks=CHUNKS, scatter_ctas=CTAS)
    sp.full_steps(2)
    sp.single_step(sp.dt_out)
    dist.barrier(); torch.cuda.synchronize()
    names, evs = [], []

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        names.append(name); evs.append(e)

    dt = sp.dt_in
    if EXCHANGE == 'p2p':
        for rep in range(2):
            mark('start')
            sp._k_junction(True, False, 0.0, True, (sp.pending_dt + dt) / 2, True); mark('k-junction + scatter to row slabs')
            sp._reduce_sums(); mark('all-reduce sums (orders the stores)')
            sp._row_pass(dt); mark('row passes + scatter to k slabs')
            sp._barrier(); mark('barrier')
        torch.cuda.synchronize()
        if rank == 0:
            for i in range(6, len(evs)):
                print(f'  {names[i]:40s} {evs[i - 1].elapsed_time(evs[i]):7.3f} ms', flush=True)
            print(f'  sub-step total                           {evs[5].elapsed_time(evs[-1]):7.3f} ms', flush=True)
        return
    mark('start')
    sp._k_junction(True, False, 0.0, True, (sp.pending_dt + dt) / 2, True); mark('k-junction (local passes)')
    sp._reduce_sums(); mark('all-reduce sums')
    sp.tp.slab_pack(sp.tbuf, sp.send, sp.nxl, sp.P, sp.nyl); mark('pack')
    sp._all_to_all(); mark('all-to-all')
    sp.rp.slab_unpack(sp.recv, sp.rbuf, sp.P, sp.nxl, sp.nyl); mark('unpack+transpose')
    sp._row_pass(dt); mark('row passes (local)')
    sp.rp.slab_pack(sp.rbuf, sp.send, sp.nyl, sp.P, sp.nxl); mark('pack')
    sp._all_to_all(); mark('all-to-all')
    sp.tp.slab_unpack(sp.recv, sp.tbuf, sp.P, sp.nyl, sp.nxl); mark('unpack+transpose')
    torch.cuda.synchronize()
    if rank == 0:
        per = sp.send.numel() * sp.send.element_size()
        for i in range(1, len(evs)):
            ms = evs[i - 1].elapsed_time(evs[i])
            extra = ''
            if names[i] == 'all-to-all':
                extra = f'  ({per * (world - 1) / world / ms * 1e-6:.0f} GB/s sent per rank)'
            elif 'pack' in names[i]:
                extra = f'  ({2 * per / ms * 1e-6:.0f} GB/s r+w)'
            print(f'  {names[i]:28s} {ms:7.3f} ms{extra}', flush=True)
        print(f'  sub-step total               {evs[0].elapsed_time(evs[-1]):7.3f} ms', flush=True)


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    global EXCHANGE, CHUNKS, CTAS
    for a in list(sys.argv):
        if a.startswith('--exchange='):
            EXCHANGE = a.split('=', 1)[1]
            sys.argv.remove(a)
        elif a.startswith('--chunks='):
            CHUNKS = int(a.split('=', 1)[1])
            sys.argv.remove(a)
        elif a.startswith('--ctas='):
            CTAS = int(a.split('=', 1)[1])
            sys.argv.remove(a)
    what = sys.argv[1] if len(sys.argv) > 1 else 'check'
    if what == 'check':
        check(rank, world, dev)
    elif what == 'breakdown':
        breakdown(rank, world, dev, int(sys.argv[2]))
    else:
        run_bench(rank, world, dev, int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 10)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
