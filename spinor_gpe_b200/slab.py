"""Slab-decomposed propagation of ONE grid over several GPUs (BASELINE config 5; SURVEY.md §8e "slab FFT").

Rank r of P owns the rows y in [r Ny/P, (r+1) Ny/P).  A split sub-step is

    k-junction on the TRANSPOSED slab  [2][Nx/P][Ny]  (contiguous y-lines): FFT_y, K factors, norm sums, iFFT_y
    all-reduce of the three norm sums (T, S0, S1)
    all-to-all transpose  -> row slab [2][Ny/P][Nx]
    row pass: iFFT_x, normalise with the GLOBAL sum, I C P C I, FFT_x
    all-to-all transpose  -> transposed slab

i.e. two all-to-alls per sub-step and k-space kept in the transposed layout, so a 2-D transform costs one
exchange.  The local passes are the hand-written kernels behind ``sgpe_pass_rows / sgpe_pass_klines /
sgpe_pass_mid / sgpe_slab_pack / sgpe_slab_unpack``; ``torch.distributed`` (NCCL over NVLink) carries the
collectives.

``exchange='p2p'`` is the fused variant: every array stays row-major (k slab = [2][Ny][Nx/P], the y-lines run down
the columns: ``sgpe_pass_kcols``), and the LAST kernel of each direction stores its output straight into the
buffers of the ranks that need it next, through peer memory mapped with CUDA IPC (NVLink / NVSwitch) — compute
and collective in one kernel, no pack / all-to-all / unpack passes.  The all-reduce of the norm sums (after the
k-junction) and a one-word all-reduce (after the row passes) order the stores before their consumers.

Lines longer than 4096 points (16384^2 of config 5) use the four-step split N = n1 * n2 inside each line:
contiguous sub-transforms (``pass_klines``) + a strided pass with the twiddles and the real-space operators
fused in (``pass_mid``): three local passes per direction instead of one.  k-space then lives in the
digit-transposed order (position k1*n2 + k2 <-> frequency k1 + n1*k2) along the split axes; the operator
tables are permuted once on the host and the state is permuted back only when it is gathered.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _capi
from ._separable import split_separable

MAGIC_GAMMA = 1 / (2 + 2 ** (1 / 3))
MAX_LINE = 4096


def four_step_split(n, forced=None):
    """n1 of the split n = n1 * n2 (1 = the line fits one CTA)."""
    if forced:
        return int(forced)
    if n <= MAX_LINE:
        return 1
    n1 = 32
    while n // n1 > MAX_LINE or n1 * n1 < n:
        n1 *= 2
    return n1


def digit_order(n, n1):
    """nat[p]: natural index held at position p of the digit-transposed order (identity when n1 == 1)."""
    if n1 == 1:
        return np.arange(n)
    n2 = n // n1
    p = np.arange(n)
    return (p // n2) + n1 * (p % n2)


def _all_ranks_are_peers(group, dev):
    """True when every rank of ``group`` sits on this host and every pair of their devices can map the other's memory
    (what sgpe_ipc_open needs).  One all-gather of (hostname, device index, peer-access row) over the group."""
    import socket
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    n_dev = torch.cuda.device_count()
    row = [bool(i == idx or torch.cuda.can_device_access_peer(idx, i)) for i in range(n_dev)]
    mine = (socket.gethostname(), int(idx), row)
    everyone = [None] * dist.get_world_size(group)
    dist.all_gather_object(everyone, mine, group=group)
    if any(host != mine[0] for host, _, _ in everyone):
        return False
    devices = [d for _, d, _ in everyone]
    if len(set(devices)) != len(devices):
        return False                      # two ranks on one device: nothing to exchange through peer memory
    return all(r[d] for _, _, r in everyone for d in devices if d < len(r))


class SeparableProblem:
    """A problem given by 1-D vectors only (no (Ny, Nx) host arrays) for grids too large to set up with
    ``PSpinor`` on the host: harmonic trap (+ optional linear detuning along y), free / Raman-shifted
    dispersion, uniform coupling; Thomas-Fermi initial state evaluated per rank on the device.  The scalar
    set-up (a_x, rescaled g_sc, chemical potential, recoil units) comes from a tiny ``PSpinor``, so it is the
    reference's arithmetic (pspinor.py:283-313, 469-501)."""

    def __init__(self, mesh_points, r_sizes, atom_num=1e4, omeg=None, g_sc=None, pop_frac=(0.5, 0.5),
                 coupling=None, wavel=790.1e-9, kin_shift=False, rot_coupling=True, detuning_slope=0.0):
        import tempfile, os
        from .pspinor import PSpinor
        with tempfile.TemporaryDirectory(prefix='sgpe_sep_') as scratch:      # PSpinor creates its data directories
            tiny = PSpinor(os.path.join(scratch, 'p') + os.sep, omeg=omeg, g_sc=g_sc,
                           mesh_points=(32, 32), r_sizes=r_sizes, atom_num=atom_num, pop_frac=pop_frac)
        self.atom_num, self.g_sc, self.chem_pot, self.pop_frac = atom_num, tiny.g_sc, tiny.chem_pot, pop_frac
        self.omeg = tiny.omeg
        nx, ny = int(mesh_points[0]), int(mesh_points[1])
        self.nx, self.ny = nx, ny
        rs = np.array(r_sizes, dtype=np.float64)
        dr = 2 * rs / np.array([nx, ny])
        self.space = {'dr': dr, 'dk': np.pi / rs, 'dv_r': float(np.prod(dr)), 'dv_k': float(np.prod(np.pi / rs)),
                      'x': np.linspace(-rs[0], rs[0], nx, endpoint=False),
                      'y': np.linspace(-rs[1], rs[1], ny, endpoint=False)}
        ks = np.pi / dr
        kx = np.linspace(-ks[0], ks[0], nx, endpoint=False)
        ky = np.linspace(-ks[1], ks[1], ny, endpoint=False)
        self.is_coupling = coupling is not None
        self.rot_coupling = rot_coupling
        self.kL_recoil, self.EL_recoil = 1.0, 1.0
        shift = 0.0
        if self.is_coupling:
            tiny.coupling_setup(wavel=wavel, kin_shift=kin_shift)
            self.kL_recoil, self.EL_recoil = tiny.kL_recoil, tiny.EL_recoil
            shift = kx * self.kL_recoil if kin_shift else 0.0
        gx = np.stack([kx ** 2 / 2 + shift, kx ** 2 / 2 - shift])
        gx = gx - gx.min(axis=1, keepdims=True)                    # kin_eng_spin - min (pspinor.py:501)
        self.kin_x, self.kin_y = gx, np.stack([ky ** 2 / 2, ky ** 2 / 2])
        y_trap = self.omeg['y'] / self.omeg['x']
        x, y = self.space['x'], self.space['y']
        det = detuning_slope * y
        self.pot_x = np.stack([x ** 2 / 2, x ** 2 / 2])
        self.pot_y = np.stack([(y_trap * y) ** 2 / 2 + det / 2, (y_trap * y) ** 2 / 2 - det / 2])
        self.omega = float(coupling) if self.is_coupling else 0.0

    def tf_rows(self, y_rows, device, dtype):
        """Thomas-Fermi amplitudes (pspinor.py:265-271) on the rows ``y_rows``: (2, len(y_rows), nx) tensor."""
        x = torch.as_tensor(self.space['x'], device=device)
        y = torch.as_tensor(np.asarray(y_rows), device=device)
        y_trap = self.omeg['y'] / self.omeg['x']
        v = (x[None, :] ** 2 + (y_trap * y[:, None]) ** 2) / 2
        prof = torch.sqrt(torch.clamp(self.chem_pot - v, min=0.0))
        g_bare = [self.g_sc['uu'], self.g_sc['dd']]
        comps = [prof * float(np.sqrt(p / abs(g))) for p, g in zip(self.pop_frac, g_bare)]
        return torch.stack(comps).to(dtype)


class SlabPropagator:
    """Distributed TensorPropagator for one trajectory.  Every rank passes the same (global) ``PSpinor``; each
    keeps only its slab on the device.  ``split_x`` / ``split_y`` force a four-step split (tests)."""

    def __init__(self, spin, t_step, time='imag', device='cuda', group=None, precision='c128', plan_kwargs=None,
                 split_x=None, split_y=None, exchange='auto', exchange_buffers=None, chunks=None, scatter_ctas=None):
        from .plan import Plan
        # group='local': ONE device, no process group — the long-line machinery (four-step lines, row-major k slab)
        # for meshes beyond 4096 points per line on a single GPU; every collective degenerates to stream order
        self.local = isinstance(group, str) and group == 'local'
        if self.local:
            self.group, self.rank, self.P = None, 0, 1
        else:
            assert dist.is_initialized(), "SlabPropagator needs an initialised torch.distributed process group"
            self.group = group
            self.rank, self.P = dist.get_rank(group), dist.get_world_size(group)
        self.dev = torch.device(device)
        assert exchange in ('auto', 'nccl', 'p2p')
        if exchange == 'auto':
            # the fused exchange stores into the other ranks' buffers (CUDA IPC): only when every rank of the group runs
            # on this node and every pair of devices has peer access; otherwise the NCCL all-to-all
            if exchange_buffers is not None or self.local:
                exchange = 'p2p'
            elif self.dev.type == 'cuda':
                exchange = 'p2p' if _all_ranks_are_peers(self.group, self.dev) else 'nccl'
            else:
                exchange = 'nccl'
        self.exchange = exchange
        self.p2p = (exchange == 'p2p')
        self._ipc = []            # (lib, own pointers, opened pointers) to release
        self.cdtype = torch.complex128 if precision == 'c128' else torch.complex64
        sep_only = isinstance(spin, SeparableProblem)
        if sep_only:
            self.nx, self.ny = spin.nx, spin.ny
            psik = None
        else:
            psik = np.array([np.asarray(p) for p in spin.psik])
            _, self.ny, self.nx = psik.shape
        P, r = self.P, self.rank
        assert self.nx % (32 * P) == 0 and self.ny % (32 * P) == 0, "mesh must split into multiples of 32 per rank"
        self.nxl, self.nyl = self.nx // P, self.ny // P
        self.time, self.dt = time, float(t_step)
        self.dt_out, self.dt_in = self.dt * MAGIC_GAMMA, self.dt * (1 - 2 * MAGIC_GAMMA)
        self.atom_num = float(spin.atom_num)
        self.dv_k = float(spin.space['dv_k'])
        self.dv_r = float(spin.space['dv_r'])
        self.n1x, self.n1y = four_step_split(self.nx, split_x), four_step_split(self.ny, split_y)
        self.nat_x, self.nat_y = digit_order(self.nx, self.n1x), digit_order(self.ny, self.n1y)
        kw = dict(plan_kwargs or {})
        ys, xs = slice(r * self.nyl, (r + 1) * self.nyl), slice(r * self.nxl, (r + 1) * self.nxl)
        self._ys, self._xs = ys, xs

        # ---- row plan: local rows (natural y), full x
        self.rp = rp = Plan(self.nx, self.nyl, 1, self.cdtype, self.dev, lines_n1=self.n1x, **kw)
        dr = spin.space['dr']
        rp.set_grid(dr[0], dr[1], spin.space['dv_r'], spin.space['dv_k'], spin.atom_num)
        rp.set_interactions(spin.g_sc['uu'], spin.g_sc['dd'], spin.g_sc['ud'])
        eiphi = None
        if spin.is_coupling and not spin.rot_coupling:
            eiphi = np.exp(1j * 2 * spin.kL_recoil * np.asarray(spin.space['x']))
        if sep_only:
            rp.set_potential_separable(spin.pot_x, np.ascontiguousarray(spin.pot_y[:, ys]))
            if spin.is_coupling and spin.omega != 0.0:
                rp.set_coupling(_capi.SGPE_COUPLING_UNIFORM, omega=np.array([spin.omega]), eiphi=eiphi)
            else:
                rp.set_coupling(_capi.SGPE_COUPLING_NONE)
        else:
            pot = np.array([np.asarray(v) for v in spin.pot_eng_spin])
            rp.set_potential(np.ascontiguousarray(pot[0, ys]), np.ascontiguousarray(pot[1, ys]))
            psep = split_separable(pot)
            if psep is not None:
                rp.set_potential_separable(psep[0], np.ascontiguousarray(psep[1][:, ys]))
            cpl = np.asarray(spin.coupling, dtype=np.float64)
            if not spin.is_coupling or not np.any(cpl):
                rp.set_coupling(_capi.SGPE_COUPLING_NONE)
                if np.any(cpl):     # eng_expect adds the coupling energy whatever is_coupling says (:319-321)
                    rp.set_energy_coupling(_capi.SGPE_COUPLING_DENSE, coupling=np.ascontiguousarray(cpl[ys]))
            elif np.all(cpl == cpl.flat[0]):
                rp.set_coupling(_capi.SGPE_COUPLING_UNIFORM, omega=np.array([cpl.flat[0]]), eiphi=eiphi)
            else:
                rp.set_coupling(_capi.SGPE_COUPLING_DENSE, coupling=np.ascontiguousarray(cpl[ys]), eiphi=eiphi)
        rp.set_time(time, self.dt)

        # ---- transposed plan: lines = local k_x positions, positions along the line = k_y positions
        self.tp = tp = Plan(self.ny, self.nxl, 1, self.cdtype, self.dev, lines_n1=self.n1y, **kw)
        tp.set_grid(dr[0], dr[1], spin.space['dv_r'], spin.space['dv_k'], spin.atom_num)
        if sep_only:
            ksep = (spin.kin_x[:, self.nat_x], spin.kin_y[:, self.nat_y])
        else:
            kin = np.array([np.asarray(k) for k in spin.kin_eng_spin])
            kin = kin[:, self.nat_y][:, :, self.nat_x]             # k-space grids in the stored (position) order
            ksep = split_separable(kin)
        if ksep is not None:      # kin[ky][kx] = gx[kx] + gy[ky]:  position table <- gy, line table <- local gx
            tp.set_kinetic_separable(np.ascontiguousarray(ksep[1]), np.ascontiguousarray(ksep[0][:, xs]))
        else:
            if self.n1y > 1:
                raise NotImplementedError("four-step lines need a separable kinetic energy grid")
            if self.p2p:
                kin_l = np.ascontiguousarray(kin[:, :, xs])                      # (2, ny, nxl) row-major k slab
            else:
                kin_l = np.ascontiguousarray(kin[:, :, xs].transpose(0, 2, 1))   # (2, nxl, ny)
            tp.set_kinetic(kin_l[0], kin_l[1])
        tp.set_time(time, self.dt)

        n_local = 2 * self.nxl * self.ny
        mk = lambda: torch.empty(n_local, dtype=self.cdtype, device=self.dev)      # noqa: E731
        if self.p2p:
            self._setup_exchange(n_local, exchange_buffers)
        else:
            self.tbuf, self.rbuf, self.send, self.recv = mk(), mk(), mk(), mk()
        self.sums = torch.zeros(4, dtype=torch.float64, device=self.dev)
        self._flag = torch.zeros(1, dtype=torch.float64, device=self.dev)
        # Chunked pipelining of the fused exchange (four-step lines only: three local passes per direction, the last
        # one NVLink-bound).  The slab is cut into `chunks` windows; the scatter pass of window c runs on a second
        # stream with a few persistent CTAs per SM while the first two passes of window c + 1 run beside it.
        if chunks is None:      # one rank has no link to hide: one window (measured: 52.5 vs 45.9 steps/s at 8192^2)
            chunks = 4 if (self.p2p and self.dev.type == 'cuda' and self.P > 1) else 1
        self.chunks_x = chunks if (self.p2p and self.n1x > 1 and self.nyl % chunks == 0) else 1
        self.chunks_y = chunks if (self.p2p and self.n1y > 1 and self.nxl % (chunks * 32) == 0) else 1
        if scatter_ctas is None:
            scatter_ctas = torch.cuda.get_device_properties(self.dev).multi_processor_count \
                if self.dev.type == 'cuda' else 3
        self.scatter_ctas = int(scatter_ctas)
        self._side = torch.cuda.Stream(self.dev) if self.dev.type == 'cuda' else None
        self._chunk_sums = torch.zeros((max(self.chunks_y, 1), 4), dtype=torch.float64, device=self.dev)
        self.mid = False
        self.pending_dt = 0.0
        self.scale_pending = False
        self.a2a_bytes = 0
        self.points = float(self.nx) * float(self.ny)
        if sep_only:
            self.set_real_space(spin.tf_rows(spin.space['y'][ys], self.dev, self.cdtype))
        else:
            self.load_psik(psik)

    # ------------------------------------------------------------------ fused exchange: buffers in peer memory
    def _setup_exchange(self, n_local, given):
        """k slab (``tbuf``, [2][Ny][Nx/P]) and row slab (``rbuf``, [2][Ny/P][Nx]) of every rank, addressable from
        this device.  Product path: cudaMalloc + CUDA IPC handles exchanged over the process group.  ``given`` (tests
        on CPU with the emulated kernels): {'k': [tensor per rank], 'r': [...]} in memory shared by the processes."""
        import ctypes
        P, r = self.P, self.rank
        if given is not None:
            self.tbuf, self.rbuf = given['k'][r].view(-1), given['r'][r].view(-1)
            assert self.tbuf.numel() == n_local and self.rbuf.numel() == n_local and self.tbuf.dtype == self.cdtype
            kptrs = [t.data_ptr() for t in given['k']]
            rptrs = [t.data_ptr() for t in given['r']]
            self._keep_given = given
        else:
            assert self.dev.type == 'cuda', "the fused exchange needs CUDA devices (or shared test buffers)"
            lib = self.rp.lib
            dev_index = self.dev.index if self.dev.index is not None else torch.cuda.current_device()
            nbytes = n_local * (16 if self.cdtype == torch.complex128 else 8)
            own, handles = [], []
            for _ in range(2):
                ptr, h = ctypes.c_void_p(), ctypes.create_string_buffer(64)
                _capi.check(lib, lib.sgpe_ipc_alloc(dev_index, nbytes, ctypes.byref(ptr), h), 'sgpe_ipc_alloc')
                own.append(ptr.value)
                handles.append(bytes(h.raw))
            everyone = [None] * P
            if P > 1:
                dist.all_gather_object(everyone, handles, group=self.group)
            kptrs, rptrs, opened = [], [], []
            for q in range(P):
                if q == r:
                    kptrs.append(own[0]); rptrs.append(own[1])
                    continue
                for which, dst in ((0, kptrs), (1, rptrs)):
                    ptr = ctypes.c_void_p()
                    _capi.check(lib, lib.sgpe_ipc_open(dev_index, everyone[q][which], ctypes.byref(ptr)), 'sgpe_ipc_open')
                    dst.append(ptr.value); opened.append(ptr.value)
            self._ipc.append((lib, own, opened))

            class _DevMem:            # raw device memory -> torch tensor (no copy)
                def __init__(self, ptr, n):
                    self.__cuda_array_interface__ = {'shape': (n,), 'typestr': '|u1', 'data': (ptr, False), 'version': 2}
            wrap = lambda ptr: torch.as_tensor(_DevMem(ptr, nbytes), device=self.dev).view(self.cdtype)   # noqa: E731
            self.tbuf, self.rbuf = wrap(own[0]), wrap(own[1])
        # row direction -> k slabs (split along x);  k direction -> row slabs (split along y)
        self.rp.set_peers(kptrs, 1, self.nxl, self.nxl, self.ny * self.nxl, r * self.nyl)
        self.tp.set_peers(rptrs, 2, self.nyl, self.nx, self.nyl * self.nx, r * self.nxl)

    def close(self):
        """Unmap / free the exchange buffers (all ranks must have stopped stepping)."""
        for lib, own, opened in self._ipc:
            self.tbuf = self.rbuf = None
            for ptr in opened:
                lib.sgpe_ipc_close(ptr)
            for ptr in own:
                lib.sgpe_ipc_free(ptr)
        self._ipc = []

    def _barrier(self):
        if self.P > 1:
            dist.all_reduce(self._flag, op=dist.ReduceOp.SUM, group=self.group)

    def _count_exchange(self):
        self.a2a_bytes += self.tbuf.numel() * self.tbuf.element_size() * (self.P - 1) // self.P

    def load_psik(self, psik):
        """Replace the state by the k-space wavefunction ``psik`` ((2, Ny, Nx), the reference's order; NumPy array or
        tensor, the same on every rank): each rank keeps its slab in the stored (digit-transposed) order."""
        if isinstance(psik, torch.Tensor):
            full = psik.reshape(2, self.ny, self.nx).to(self.dev)
            stored = full.index_select(1, torch.as_tensor(self.nat_y, device=self.dev)) \
                         .index_select(2, torch.as_tensor(self.nat_x, device=self.dev))
            local = stored[:, :, self._xs] if self.p2p else stored[:, :, self._xs].transpose(1, 2)
            local = local.contiguous().reshape(-1).to(self.cdtype)
        else:
            stored = np.asarray(psik).reshape(2, self.ny, self.nx)[:, self.nat_y][:, :, self.nat_x]
            if self.p2p:
                local = np.ascontiguousarray(stored[:, :, self._xs])                  # (2, ny, nxl) row-major k slab
            else:
                local = np.ascontiguousarray(stored[:, :, self._xs].transpose(0, 2, 1))   # (2, nxl, ny) transposed slab
            local = torch.as_tensor(local).reshape(-1).to(self.cdtype)
        if self.p2p:
            self._barrier()           # the peers are done with whatever they were storing here
        self.tbuf.copy_(local)
        self.mid, self.scale_pending, self.pending_dt = False, False, 0.0
        if self.p2p:
            self._barrier()           # nobody stores into a peer that is still loading

    def set_real_space(self, psi_rows):
        """Load a REAL-space state given by this rank's rows, (2, Ny/P, Nx) on the device: the distributed forward
        transform (x-lines, transpose, y-lines) leaves it in the stored k-space layout.  The overall scale is
        irrelevant (the first sub-step renormalises to the atom number)."""
        ny0 = self._ys.start
        sign = 1.0 - 2.0 * ((torch.arange(self.nx, device=self.dev)[None, :]
                             + torch.arange(ny0, ny0 + self.nyl, device=self.dev)[:, None]) % 2)
        if self.p2p:
            self._barrier()           # the peers are done with whatever they were reading
        self.rbuf.view(2, self.nyl, self.nx).copy_(psi_rows * sign.to(psi_rows.real.dtype))    # the fftshift sign
        if self.n1x > 1:
            self.rp.pass_mid(self.rbuf, False, False, False, 0.0, True, True, None, 0.0)
        self.rp.pass_klines(self.rbuf, True, False, 0.0, False, 0.0, False, None, scatter=self.p2p)
        if self.p2p:
            self._count_exchange()
            self._barrier()
            if self.n1y > 1:
                self.tp.pass_mid(self.tbuf, False, False, False, 0.0, True, True, None, 0.0, inner=self.nxl)
            self.tp.pass_kcols(self.tbuf, True, False, 0.0, False, 0.0, False, None)
            self._barrier()
        else:
            self._to_lines()
            if self.n1y > 1:
                self.tp.pass_mid(self.tbuf, False, False, False, 0.0, True, True, None, 0.0)
            self.tp.pass_klines(self.tbuf, True, False, 0.0, False, 0.0, False, None)
        self.mid, self.scale_pending = False, False

    # ------------------------------------------------------------------ collectives
    def _all_to_all(self):
        s, r = torch.view_as_real(self.send), torch.view_as_real(self.recv)
        if self.P > 1:
            dist.all_to_all_single(r.view(-1), s.view(-1), group=self.group)
        else:
            r.copy_(s)
        self.a2a_bytes += self.send.numel() * self.send.element_size() * (self.P - 1) // self.P

    def _to_rows(self):
        self.tp.slab_pack(self.tbuf, self.send, self.nxl, self.P, self.nyl)      # [2][nxl][P*nyl] -> [P][2][nxl][nyl]
        self._all_to_all()
        self.rp.slab_unpack(self.recv, self.rbuf, self.P, self.nxl, self.nyl)    # -> [2][nyl][P*nxl]

    def _to_lines(self):
        self.rp.slab_pack(self.rbuf, self.send, self.nyl, self.P, self.nxl)      # [2][nyl][P*nxl] -> [P][2][nyl][nxl]
        self._all_to_all()
        self.tp.slab_unpack(self.recv, self.tbuf, self.P, self.nyl, self.nxl)    # -> [2][nxl][P*nyl]

    def _reduce_sums(self):
        if self.P > 1:
            dist.all_reduce(self.sums, op=dist.ReduceOp.SUM, group=self.group)

    # ------------------------------------------------------------------ second stream of the chunked exchange
    def _fork(self):
        """The side stream waits for everything enqueued on the current stream so far."""
        if self._side is not None:
            self._side.wait_stream(torch.cuda.current_stream(self.dev))

    def _join(self):
        if self._side is not None:
            torch.cuda.current_stream(self.dev).wait_stream(self._side)

    def _on_side(self):
        import contextlib
        return torch.cuda.stream(self._side) if self._side is not None else contextlib.nullcontext()

    # ------------------------------------------------------------------ local passes (one or three per direction)
    def _k_junction(self, do_fwd, has_a, tau_a, has_b, tau_b, do_inv):
        tp = self.tp
        if self.p2p and self.chunks_y > 1 and do_inv:
            cw = self.nxl // self.chunks_y
            for c in range(self.chunks_y):
                tp.window(c * cw, cw, c, 0)
                if do_fwd:
                    tp.pass_mid(self.tbuf, False, False, False, 0.0, True, True, None, 0.0, inner=self.nxl)
                tp.pass_kcols(self.tbuf, do_fwd, has_a, tau_a, has_b, tau_b, do_inv, self._chunk_sums[c])
                self._fork()
                with self._on_side():      # NVLink-bound: a few persistent CTAs per SM, beside the next window's passes
                    tp.window(c * cw, cw, c, self.scatter_ctas)
                    tp.pass_mid(self.tbuf, True, True, False, 0.0, False, False, None, 0.0, inner=self.nxl, scatter=True)
            tp.window()
            self._join()
            self.sums.copy_(self._chunk_sums.sum(0))
            self._count_exchange()
            return
        if self.p2p:          # row-major k slab; the kernel that finishes the inverse stores into the peers' row slabs
            if self.n1y == 1:
                tp.pass_kcols(self.tbuf, do_fwd, has_a, tau_a, has_b, tau_b, do_inv, self.sums, scatter=do_inv)
            else:
                if do_fwd:
                    tp.pass_mid(self.tbuf, False, False, False, 0.0, True, True, None, 0.0, inner=self.nxl)
                tp.pass_kcols(self.tbuf, do_fwd, has_a, tau_a, has_b, tau_b, do_inv, self.sums)
                if do_inv:
                    tp.pass_mid(self.tbuf, True, True, False, 0.0, False, False, None, 0.0, inner=self.nxl, scatter=True)
            if do_inv:
                self._count_exchange()
            return
        if self.n1y == 1:
            tp.pass_klines(self.tbuf, do_fwd, has_a, tau_a, has_b, tau_b, do_inv, self.sums)
            return
        if do_fwd:        # strided forward over y1, then the four-step twiddle
            tp.pass_mid(self.tbuf, False, False, False, 0.0, True, True, None, 0.0)
        tp.pass_klines(self.tbuf, do_fwd, has_a, tau_a, has_b, tau_b, do_inv, self.sums)
        if do_inv:        # conjugate twiddle, strided inverse over k1
            tp.pass_mid(self.tbuf, True, True, False, 0.0, False, False, None, 0.0)

    def _row_pass(self, dt_sub):
        rp = self.rp
        if self.p2p and self.chunks_x > 1:
            rw = self.nyl // self.chunks_x
            for c in range(self.chunks_x):
                rp.window(c * rw, rw, c, 0)
                rp.pass_klines(self.rbuf, False, False, 0.0, False, 0.0, True, None)
                rp.pass_mid(self.rbuf, True, True, True, dt_sub, True, True, self.sums, self.points)
                self._fork()
                with self._on_side():
                    rp.window(c * rw, rw, c, self.scatter_ctas)
                    rp.pass_klines(self.rbuf, True, False, 0.0, False, 0.0, False, None, scatter=True)
            rp.window()
            self._join()
        elif self.n1x == 1:
            rp.pass_rows(self.rbuf, dt_sub, self.sums, self.points, scatter=self.p2p)
        else:
            rp.pass_klines(self.rbuf, False, False, 0.0, False, 0.0, True, None)       # contiguous inverse over k2
            rp.pass_mid(self.rbuf, True, True, True, dt_sub, True, True, self.sums, self.points)
            rp.pass_klines(self.rbuf, True, False, 0.0, False, 0.0, False, None, scatter=self.p2p)   # forward over n2
        if self.p2p:
            self._count_exchange()

    # ------------------------------------------------------------------ stepping
    def single_step(self, dt_sub, pops_out=None):
        imag = (self.time == 'imag')
        if not self.mid:
            self._k_junction(False, False, 0.0, True, dt_sub / 2, True)
        elif pops_out is not None and imag:
            self._k_junction(True, True, self.pending_dt / 2, True, dt_sub / 2, True)
        else:
            self._k_junction(True, False, 0.0, True, (self.pending_dt + dt_sub) / 2, True)
        self._reduce_sums()
        if pops_out is not None and self.mid:
            pops_out.copy_(self.atom_num * self.sums[1:3] / (self.sums[1] + self.sums[2]))
        if self.p2p:          # the all-reduce above ordered the peers' stores into rbuf before this point
            self._row_pass(dt_sub)
            self._barrier()
        else:
            self._to_rows()
            self._row_pass(dt_sub)
            self._to_lines()
        self.mid, self.pending_dt, self.scale_pending = True, dt_sub, False

    def close_junction(self, pops_out=None):
        if not self.mid:
            return
        self._k_junction(True, True, self.pending_dt / 2, False, 0.0, False)
        self._reduce_sums()
        if pops_out is not None:
            pops_out.copy_(self.atom_num * self.sums[1:3] / (self.sums[1] + self.sums[2]))
        self.mid, self.scale_pending = False, True

    def full_steps(self, n, pops=None):
        """n full steps; pops: optional (n, 2) float64 tensor on the device (same on every rank)."""
        pending = None
        for i in range(n):
            self.single_step(self.dt_out, pending)
            self.single_step(self.dt_in)
            self.single_step(self.dt_out)
            pending = pops[i] if pops is not None else None
        if n > 0:
            self.close_junction(pending)

    def local_psik(self):
        """Normalised k-space state of this rank in the transposed (and, for split axes, digit-transposed)
        layout, (2, Nx/P, Ny)."""
        self.close_junction()
        if self.p2p:
            out = self.tbuf.view(2, self.ny, self.nxl).transpose(1, 2)
        else:
            out = self.tbuf.view(2, self.nxl, self.ny)
        if self.scale_pending:
            scale = torch.sqrt(self.atom_num / (self.dv_k * (self.sums[1] + self.sums[2])))
            out = out * scale.to(out.real.dtype)
        return out

    def real_space_rows(self):
        """This rank's rows of the normalised REAL-space state, (2, Ny/P, Nx) on the device — ttools.ifft_2d of the
        current k-space state (tensor_tools.py:248-256) by the distributed inverse transform: y-lines on the k slab,
        exchange, x-lines on the row slab.  The k-space state is left as it is (the inverse runs on a copy); the row
        slab buffer, free between sub-steps, receives the result.  Fused-exchange layout only."""
        if not self.p2p:
            raise NotImplementedError("real_space_rows needs the row-major k slab (exchange='p2p')")
        self.close_junction()
        keep = self.tbuf.clone()
        # norm of the state being transformed: the atom number when the lazy normalisation is pending (that is what
        # materialising it gives), else whatever was loaded (the reference's ifft_2d does not renormalise)
        if self.scale_pending:
            norm_k = torch.full((1,), self.atom_num, dtype=torch.float64, device=self.dev)
        else:
            norm_k = (keep.real.double() ** 2 + keep.imag.double() ** 2).sum().reshape(1) * self.dv_k
            if self.P > 1:
                dist.all_reduce(norm_k, op=dist.ReduceOp.SUM, group=self.group)
        tp, rp = self.tp, self.rp
        if self.n1y == 1:
            tp.pass_kcols(self.tbuf, False, False, 0.0, False, 0.0, True, None, scatter=True)
        else:
            tp.pass_kcols(self.tbuf, False, False, 0.0, False, 0.0, True, None)
            tp.pass_mid(self.tbuf, True, True, False, 0.0, False, False, None, 0.0, inner=self.nxl, scatter=True)
        self._count_exchange()
        self._barrier()
        rp.pass_klines(self.rbuf, False, False, 0.0, False, 0.0, True, None)           # inverse over x (or over k2)
        if self.n1x > 1:
            rp.pass_mid(self.rbuf, True, True, False, 0.0, False, False, None, 0.0)    # conj twiddle, inverse over k1
        self.tbuf.copy_(keep)
        del keep
        ny0 = self._ys.start
        sign = 1.0 - 2.0 * ((torch.arange(self.nx, device=self.dev)[None, :]
                             + torch.arange(ny0, ny0 + self.nyl, device=self.dev)[:, None]) % 2)
        rows = self.rbuf.view(2, self.nyl, self.nx)
        psi = rows * sign.to(rows.real.dtype)                   # the ifftshift of the reference is this sign
        # scale by Parseval (dv_r sum|psi|^2 = dv_k sum|psi_k|^2, the invariant of the reference's transforms),
        # whatever factors the individual passes carry
        tot = (psi.real.double() ** 2 + psi.imag.double() ** 2).sum().reshape(1)
        if self.P > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=self.group)
        psi = psi * torch.sqrt(norm_k[0] / (self.dv_r * tot[0])).to(rows.real.dtype)
        self._barrier()
        return psi

    def gather_psik(self):
        """Full (2, Ny, Nx) k-space state in the reference's order on every rank (tests / small grids only)."""
        loc = torch.view_as_real(self.local_psik().contiguous())
        if self.P > 1:
            parts = [torch.empty_like(loc) for _ in range(self.P)]
            dist.all_gather(parts, loc, group=self.group)
        else:
            parts = [loc]
        full = torch.cat([torch.view_as_complex(p) for p in parts], dim=1)       # (2, Nx, Ny), stored order
        full = full.transpose(1, 2).contiguous()
        inv_x = torch.as_tensor(np.argsort(self.nat_x), device=full.device)
        inv_y = torch.as_tensor(np.argsort(self.nat_y), device=full.device)
        return full.index_select(2, inv_x).index_select(1, inv_y).contiguous()


class LongLinePlan:
    """The subset of ``plan.Plan`` that ``TensorPropagator`` drives, for ONE trajectory on ONE GPU whose mesh has more
    than 4096 points along a line (8192^2, 16384^2 in 180 GB of HBM): a ``SlabPropagator`` with one rank and no
    process group.  Lines are split four-step style (contiguous sub-lines + a strided pass with the twiddles and the
    real-space operators fused in), k-space is kept in the digit-transposed order and un-permuted only when the
    state is read.  Needs a separable kinetic energy grid (everything ``PSpinor`` builds)."""

    def __init__(self, spin, t_step, time, device, precision='c128', **slab_kwargs):
        self.sp = SlabPropagator(spin, t_step, time=time, device=device, group='local', precision=precision,
                                 exchange='p2p', **slab_kwargs)
        self.nx, self.ny, self.batch = self.sp.nx, self.sp.ny, 1
        self.device = self.sp.dev
        self.cdtype = self.sp.cdtype
        self._kl_keep = None

    def load(self, psik):
        self.sp.load_psik(psik if isinstance(psik, torch.Tensor) else np.asarray(psik))

    def store(self, out=None):
        full = self.sp.gather_psik().reshape(1, 2, self.ny, self.nx)
        if out is not None:
            out.copy_(full)
            return out
        return full

    def substeps(self):
        return self.sp.dt_out, self.sp.dt_in

    def single_step(self, dt_sub):
        self.sp.single_step(float(dt_sub))

    def full_steps(self, n, pops=None, first=0, energy=None, kl_term=0.0, unwrap='none'):
        """As ``Plan.full_steps``.  With ``energy`` (float64 (1, n_total, 4)) the energy expectation of every step
        boundary is evaluated too — here through the stand-alone evaluation after each step (the fused side chain of
        the short-line kernels does not exist for four-step lines)."""
        if energy is None:
            self.sp.full_steps(int(n), None if pops is None else pops[0, first:first + n])
            return
        for i in range(int(n)):
            self.sp.full_steps(1, None if pops is None else pops[0, first + i:first + i + 1])
            energy[0, first + i].copy_(self.energy(None, kl_term=kl_term, unwrap=unwrap)[0])

    def kinetic_spectral(self, psik=None, kin_x=None, kin_y=None):
        """dv_k * sum_k kin_c |psi_k,c|^2 per component from the separable kinetic grid kin_c = kin_x[c][kx] +
        kin_y[c][ky], (1, 2) float64: two marginal sums of the k-space density on the device."""
        full = self.store() if psik is None else psik.reshape(1, 2, self.ny, self.nx).to(self.device)
        dens = full[0].real.double() ** 2 + full[0].imag.double() ** 2            # (2, ny, nx)
        kx = torch.as_tensor(kin_x, dtype=torch.float64, device=self.device)
        ky = torch.as_tensor(kin_y, dtype=torch.float64, device=self.device)
        out = (dens.sum(1) * kx).sum(1) + (dens.sum(2) * ky).sum(1)
        return (out * self.sp.dv_k).reshape(1, 2)

    def real_space(self):
        """(1, 2, ny, nx) normalised real-space state (ttools.ifft_2d of the current state)."""
        return self.sp.real_space_rows().contiguous().reshape(1, 2, self.ny, self.nx)

    def energy(self, psik=None, kl_term=0.0, unwrap='none'):
        sp = self.sp
        saved = None
        if psik is not None:          # evaluate another state: park the current one
            sp.close_junction()
            saved = (sp.tbuf.clone(), sp.scale_pending, sp.sums.clone())
            sp.load_psik(psik)
        psi = self.real_space()
        out = sp.rp.energy_real_space(psi, kl_term=kl_term, unwrap=unwrap)
        if saved is not None:
            sp.tbuf.copy_(saved[0])
            sp.scale_pending = saved[1]
            sp.sums.copy_(saved[2])
        return out

    def set_option(self, name, value):
        self.sp.rp.set_option(name, value)
        self.sp.tp.set_option(name, value)

    def launch_count(self):
        return self.sp.rp.launch_count() + self.sp.tp.launch_count()

    def close(self):
        self.sp.close()
        self.sp.rp.close()
        self.sp.tp.close()
