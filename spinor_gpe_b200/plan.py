"""``Plan`` — torch-tensor level wrapper of one ``sgpe_plan`` (the C ABI of include/sgpe.h).

PyTorch is used for device memory and streams only; all arithmetic happens in libsgpe.so."""
import ctypes

import numpy as np
import torch

from . import _capi
from ._lib import lib, require_cuda


# TensorPropagator.eng_expect(unwrap=...): 'none' the wrapped phase as is, 'local' locally wrapped differences,
# 'herraez' the reference's reliability-sorted unwrapping (sgpe_unwrap_phase)
UNWRAP_MODES = {'none': 0, 'local': 1, 'herraez': 2}


# line lengths the kernels transform: powers of two from 32 to 4096 in the fused register-resident passes (one CTA per
# line); any other EVEN length up to 4096 whose prime factors are 2, 3, 5, 7 in the generic passes (csrc/generic.cuh);
# longer power-of-two lines as a four-step split n1 x n2 of two fused lengths (slab.LongLinePlan)
MIN_LINE, MAX_LINE, MAX_LONG_LINE = 2, 4096, 4096 * 1024


def _smooth(n):
    for q in (2, 3, 5, 7):
        while n % q == 0:
            n //= q
    return n == 1


def supported_mesh(n):
    """(ok, reason) for a mesh size along one axis.  The reference asserts even sizes only (pspinor.py:331-332, its
    message asks for powers of two) and leaves the rest to cuFFT / MKL."""
    n = int(n)
    if n < MIN_LINE or n % 2:
        return False, "the number of mesh points must be even (pspinor.py:331-332)"
    if n <= MAX_LINE:
        if _smooth(n):
            return True, ''
        return False, "the in-house FFT takes sizes whose prime factors are 2, 3, 5 and 7"
    if n & (n - 1):
        return False, f"lines longer than {MAX_LINE} points must be powers of two (four-step split)"
    if n > MAX_LONG_LINE:
        return False, f"lines longer than {MAX_LONG_LINE} points are not supported"
    return True, ''


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _dp(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class Plan:
    """One propagator plan on one GPU for ``batch`` independent trajectories of an (ny, nx) mesh."""

    def __init__(self, nx, ny, batch=1, dtype=torch.complex128, device='cuda', _lib=None, lines_n1=None):
        # ``_lib`` is a test hook: tests/emu_harness.py passes the emulated build of the same sources so that
        # host-side logic (sharding, slab orchestration) can run on CPU tensors; the product never sets it.
        self.device = torch.device(device)
        if _lib is None:
            require_cuda()
            self.lib = lib()
            if self.device.type != 'cuda':
                raise ValueError("spinor_gpe_b200 plans live on CUDA devices only (no CPU fallback)")
            if self.device.index is None:
                self.device = torch.device('cuda', torch.cuda.current_device())
        else:
            self.lib = _lib
        self.nx, self.ny, self.batch = int(nx), int(ny), int(batch)
        self.cdtype = dtype
        self.rdtype = torch.float64
        code = _capi.SGPE_C128 if dtype == torch.complex128 else _capi.SGPE_C64
        self.h = ctypes.c_void_p()
        if lines_n1 is None:
            self._chk(self.lib.sgpe_plan_create(ctypes.byref(self.h), self.nx, self.ny, self.batch, code,
                                                self.device.index or 0), 'sgpe_plan_create')
        else:
            # line plan of the slab mode: ny lines of nx points, caller-owned buffers, optional four-step split
            self._chk(self.lib.sgpe_plan_create_lines(ctypes.byref(self.h), self.nx, self.ny, int(lines_n1), code,
                                                      self.device.index or 0), 'sgpe_plan_create_lines')
        self.keep = {}

    # ------------------------------------------------------------------ plumbing
    def _chk(self, rc, what):
        _capi.check(self.lib, rc, what)

    def close(self):
        if getattr(self, 'h', None):
            self.lib.sgpe_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _f64(self, key, arr):
        t = torch.as_tensor(np.asarray(arr) if not isinstance(arr, torch.Tensor) else arr)
        t = t.to(device=self.device, dtype=torch.float64).contiguous()
        self.keep[key] = t
        return t

    def _state(self, t):
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(np.asarray(t))
        t = t.to(device=self.device, dtype=self.cdtype).contiguous()
        if t.numel() != self.batch * 2 * self.ny * self.nx:
            raise ValueError(f"state has {t.numel()} elements, plan needs "
                             f"{self.batch}x2x{self.ny}x{self.nx}")
        return t

    def new_state(self):
        return torch.empty((self.batch, 2, self.ny, self.nx), dtype=self.cdtype, device=self.device)

    @property
    def stream(self):
        if self.device.type != 'cuda':
            return None
        return _stream_ptr(self.device)

    # ------------------------------------------------------------------ problem definition
    def set_grid(self, dx, dy, dv_r, dv_k, atom_num):
        self._chk(self.lib.sgpe_set_grid(self.h, float(dx), float(dy), float(dv_r), float(dv_k),
                                         float(atom_num)), 'sgpe_set_grid')

    def set_interactions(self, g_uu, g_dd, g_ud):
        self._chk(self.lib.sgpe_set_interactions(self.h, float(g_uu), float(g_dd), float(g_ud)),
                  'sgpe_set_interactions')

    def set_kinetic(self, kin0, kin1, batched=False):
        """kin_c: (ny,nx) or, batched, (B,ny,nx) float64."""
        k = self._f64('kin', torch.stack([torch.as_tensor(np.asarray(kin0)) if not isinstance(kin0, torch.Tensor) else kin0,
                                          torch.as_tensor(np.asarray(kin1)) if not isinstance(kin1, torch.Tensor) else kin1]))
        plane = self.nx * self.ny
        per = k[0].numel()
        self._chk(self.lib.sgpe_set_kinetic(self.h, _dp(k[0]), ctypes.c_void_p(k.data_ptr() + 8 * per),
                                            plane if batched else 0), 'sgpe_set_kinetic')

    def set_potential(self, pot0, pot1, batched=False, shared=False):
        p0 = self._f64('pot0', pot0)
        p1 = p0 if shared else self._f64('pot1', pot1)
        self._chk(self.lib.sgpe_set_potential(self.h, _dp(p0), _dp(p1),
                                              self.nx * self.ny if batched else 0), 'sgpe_set_potential')

    def set_kinetic_separable(self, kin_x, kin_y, batched=False):
        """kin_c[ky][kx] = kin_x[c][kx] + kin_y[c][ky]; kin_x (2,nx) / kin_y (2,ny) or, batched, (B,2,n)."""
        kx, ky = self._f64('kin_x', kin_x), self._f64('kin_y', kin_y)
        self._chk(self.lib.sgpe_set_kinetic_separable(self.h, _dp(kx), _dp(ky), 2 * self.nx if batched else 0,
                                                      2 * self.ny if batched else 0), 'sgpe_set_kinetic_separable')

    def set_potential_separable(self, pot_x, pot_y, batched=False):
        px, py = self._f64('pot_x', pot_x), self._f64('pot_y', pot_y)
        self._chk(self.lib.sgpe_set_potential_separable(self.h, _dp(px), _dp(py), 2 * self.nx if batched else 0,
                                                        2 * self.ny if batched else 0), 'sgpe_set_potential_separable')

    def set_coupling(self, mode, coupling=None, omega=None, eiphi=None, batched=False):
        c = self._f64('coupling', coupling) if coupling is not None else None
        o = self._f64('omega', omega) if omega is not None else None
        e = None
        if eiphi is not None:
            e = torch.as_tensor(np.asarray(eiphi) if not isinstance(eiphi, torch.Tensor) else eiphi)
            e = e.to(device=self.device, dtype=self.cdtype).contiguous()
            self.keep['eiphi'] = e
        self._chk(self.lib.sgpe_set_coupling(self.h, int(mode), _dp(c), self.nx * self.ny if batched else 0,
                                             _dp(o), _dp(e)), 'sgpe_set_coupling')

    def set_energy_coupling(self, mode, coupling=None, omega=None, batched=False):
        """Coupling grid of ``energy()`` when it is not what the stepping applies (the reference's eng_expect always
        uses ``self.coupling``, tensor_propagator.py:319-321); mode -1 = follow ``set_coupling``."""
        c = self._f64('e_coupling', coupling) if coupling is not None else None
        o = self._f64('e_omega', omega) if omega is not None else None
        self._chk(self.lib.sgpe_set_energy_coupling(self.h, int(mode), _dp(c), self.nx * self.ny if batched else 0,
                                                    _dp(o)), 'sgpe_set_energy_coupling')

    def set_option(self, name, value):
        self._chk(self.lib.sgpe_set_option(self.h, name.encode(), int(value)), 'sgpe_set_option')

    def set_time(self, mode, dt):
        code = _capi.SGPE_TIME_IMAG if mode == 'imag' else _capi.SGPE_TIME_REAL
        self._chk(self.lib.sgpe_set_time(self.h, code, float(dt)), 'sgpe_set_time')

    def substeps(self):
        a, b = ctypes.c_double(), ctypes.c_double()
        self._chk(self.lib.sgpe_substeps(self.h, ctypes.byref(a), ctypes.byref(b)), 'sgpe_substeps')
        return a.value, b.value

    # ------------------------------------------------------------------ state and stepping
    def load(self, psik):
        t = self._state(psik)
        self._chk(self.lib.sgpe_load_psik(self.h, _dp(t), self.stream), 'sgpe_load_psik')

    def store(self, out=None):
        if out is None:
            out = self.new_state()
        self._chk(self.lib.sgpe_store_psik(self.h, _dp(out), self.stream), 'sgpe_store_psik')
        return out

    def single_step(self, dt_sub):
        self._chk(self.lib.sgpe_single_step(self.h, float(dt_sub), self.stream), 'sgpe_single_step')

    def full_steps(self, n, pops=None, first=0, energy=None, kl_term=0.0, unwrap='none'):
        """pops: optional float64 CUDA tensor (B, n_total, 2); step i writes row first+i.  energy: optional float64
        CUDA tensor (B, n_total, 4) — the energy expectation [E_total, E_kin, E_pot, E_int] of every step boundary
        (row first+i): ``unwrap`` 'none' or 'local' evaluated behind the junction passes without synchronisation,
        'herraez' (the reference's definition) through the stand-alone evaluation of every closed step."""
        stride = 0
        if pops is not None:
            assert pops.is_cuda and pops.dtype == torch.float64 and pops.is_contiguous()
            stride = pops.shape[1] * 2
        if energy is None:
            self._chk(self.lib.sgpe_full_steps(self.h, int(n), _dp(pops), stride, int(first), self.stream),
                      'sgpe_full_steps')
            return
        assert energy.is_cuda and energy.dtype == torch.float64 and energy.is_contiguous() and energy.shape[-1] == 4
        self._chk(self.lib.sgpe_full_steps_energy(self.h, int(n), _dp(pops), stride, int(first), _dp(energy),
                                                  energy.shape[1] * 4, int(first), UNWRAP_MODES[unwrap], float(kl_term),
                                                  self.stream), 'sgpe_full_steps_energy')

    def run_host(self, psik_host, n_steps, want_pops=True, out=None, pops=None):
        """Host-buffer path (H2D + steps + D2H inside the call).  psik_host: CPU tensor/ndarray.  ``out`` / ``pops``:
        optional caller-owned (ideally pinned) result buffers of the right shape, reused across calls."""
        a = torch.as_tensor(np.asarray(psik_host) if not isinstance(psik_host, torch.Tensor) else psik_host)
        a = a.to(dtype=self.cdtype).contiguous()
        if out is None:
            out = torch.empty_like(a, pin_memory=True)
        elif out.dtype != a.dtype or out.numel() != a.numel() or not out.is_contiguous():
            raise ValueError("out must be a contiguous buffer of the state's dtype and size")
        if want_pops and pops is None:
            pops = torch.zeros((self.batch, n_steps, 2), dtype=torch.float64, pin_memory=True)
        elif want_pops and (pops.dtype != torch.float64 or pops.numel() != self.batch * n_steps * 2):
            raise ValueError("pops must be a float64 buffer of (batch, n_steps, 2)")
        elif not want_pops:
            pops = None
        self._chk(self.lib.sgpe_run_host(self.h, _dp(a), _dp(out), int(n_steps), _dp(pops), self.stream),
                  'sgpe_run_host')
        return out, pops

    # ------------------------------------------------------------------ helpers of tensor_tools
    def fft2d(self, t, inverse=False, out=None):
        t = self._state(t)
        out = torch.empty_like(t) if out is None else out
        self._chk(self.lib.sgpe_fft2d(self.h, _dp(t), _dp(out), int(inverse), self.stream), 'sgpe_fft2d')
        return out

    def fft1d(self, t, axis, inverse=False):
        t = self._state(t)
        out = torch.empty_like(t)
        self._chk(self.lib.sgpe_fft1d(self.h, _dp(t), _dp(out), int(axis), int(inverse), self.stream), 'sgpe_fft1d')
        return out

    def sumsq(self, t):
        t = self._state(t)
        out = torch.zeros((self.batch, 2), dtype=torch.float64, device=self.device)
        self._chk(self.lib.sgpe_sumsq(self.h, _dp(t), _dp(out), self.stream), 'sgpe_sumsq')
        return out

    def normalise(self, t, vol):
        t = self._state(t)
        out = torch.empty_like(t)
        self._chk(self.lib.sgpe_normalise(self.h, _dp(t), _dp(out), float(vol), self.stream), 'sgpe_normalise')
        return out

    def energy(self, psik=None, kl_term=0.0, unwrap='none'):
        """[E_total, E_kin, E_pot, E_int] per trajectory, (B,4) float64 CUDA tensor."""
        t = self._state(psik) if psik is not None else None
        out = torch.zeros((self.batch, 4), dtype=torch.float64, device=self.device)
        mode = UNWRAP_MODES[unwrap]
        self._chk(self.lib.sgpe_energy(self.h, _dp(t), mode, float(kl_term), _dp(out), self.stream), 'sgpe_energy')
        return out

    def kinetic_spectral(self, psik=None):
        """dv_k * sum_k kin_c |psi_k,c|^2 per trajectory and component, (B, 2) float64 CUDA tensor."""
        t = self._state(psik) if psik is not None else None
        out = torch.zeros((self.batch, 2), dtype=torch.float64, device=self.device)
        self._chk(self.lib.sgpe_kinetic_spectral(self.h, _dp(t), _dp(out), self.stream), 'sgpe_kinetic_spectral')
        return out

    def gradient(self, field, h0, h1):
        """np.gradient(field, h0, h1) of ONE (ny, nx) field on the device (real or complex, converted to the plan's
        precision): [d/d(axis 0), d/d(axis 1)] as two tensors of the field's shape."""
        is_c = field.is_complex()
        rdtype = torch.float64 if self.cdtype == torch.complex128 else torch.float32
        f = field.to(device=self.device, dtype=self.cdtype if is_c else rdtype).contiguous()
        if tuple(f.shape) != (self.ny, self.nx):
            raise ValueError(f"field of {tuple(f.shape)}, plan has ({self.ny}, {self.nx})")
        g0, g1 = torch.empty_like(f), torch.empty_like(f)
        self._chk(self.lib.sgpe_gradient(self.h, _dp(f), int(is_c), float(h0), float(h1), _dp(g0), _dp(g1), self.stream),
                  'sgpe_gradient')
        return [g0, g1]

    def energy_real_space(self, psi, kl_term=0.0, unwrap='none'):
        """The energy functional on a real-space state (B, 2, ny, nx) already on the device; also valid on line
        plans (meshes beyond 4096 points per line)."""
        if psi.dtype != self.cdtype or not psi.is_contiguous() or psi.numel() != self.batch * 2 * self.ny * self.nx:
            raise ValueError("psi must be a contiguous (B, 2, ny, nx) tensor of the plan's dtype")
        out = torch.zeros((self.batch, 4), dtype=torch.float64, device=self.device)
        self._chk(self.lib.sgpe_energy_real_space(self.h, _dp(psi), UNWRAP_MODES[unwrap], float(kl_term), _dp(out),
                                                  self.stream), 'sgpe_energy_real_space')
        return out

    def unwrap_phase(self, field, mask=False):
        """Unwrapped phase of every (ny, nx) plane of ``field`` — a complex tensor (phase = angle) or a float64
        tensor of wrapped angles; float64 CUDA tensor of the same shape (ttools.phase(..., uwrap=True))."""
        if not isinstance(field, torch.Tensor):
            field = torch.as_tensor(np.asarray(field))
        kind = 0 if field.is_complex() else 1
        field = field.to(device=self.device, dtype=self.cdtype if kind == 0 else torch.float64).contiguous()
        if field.shape[-2:] != (self.ny, self.nx):
            raise ValueError(f"planes of {tuple(field.shape[-2:])}, plan has ({self.ny}, {self.nx})")
        out = torch.empty(field.shape, dtype=torch.float64, device=self.device)
        nplanes = field.numel() // (self.ny * self.nx)
        self._chk(self.lib.sgpe_unwrap_phase(self.h, _dp(field), kind, nplanes, int(bool(mask)), _dp(out), self.stream),
                  'sgpe_unwrap_phase')
        return out

    # ------------------------------------------------------------------ slab-mode local passes
    def pass_rows(self, buf, dt_sub, totals, global_points, scatter=False):
        self._chk(self.lib.sgpe_pass_rows(self.h, _dp(buf), float(dt_sub), _dp(totals), float(global_points),
                                          int(scatter), self.stream), 'sgpe_pass_rows')

    def pass_klines(self, buf, do_fwd, has_a, tau_a, has_b, tau_b, do_inv, sums, scatter=False):
        self._chk(self.lib.sgpe_pass_klines(self.h, _dp(buf), int(do_fwd), int(has_a), float(tau_a), int(has_b),
                                            float(tau_b), int(do_inv), _dp(sums), int(scatter), self.stream),
                  'sgpe_pass_klines')

    def pass_kcols(self, buf, do_fwd, has_a, tau_a, has_b, tau_b, do_inv, sums, scatter=False):
        self._chk(self.lib.sgpe_pass_kcols(self.h, _dp(buf), int(do_fwd), int(has_a), float(tau_a), int(has_b),
                                           float(tau_b), int(do_inv), _dp(sums), int(scatter), self.stream),
                  'sgpe_pass_kcols')

    def pass_mid(self, buf, pre_tw, do_inv, do_pw, dt_sub, do_fwd, post_tw, totals, global_points, inner=1,
                 scatter=False):
        self._chk(self.lib.sgpe_pass_mid(self.h, _dp(buf), int(pre_tw), int(do_inv), int(do_pw), float(dt_sub),
                                         int(do_fwd), int(post_tw), _dp(totals), float(global_points), int(inner),
                                         int(scatter), self.stream), 'sgpe_pass_mid')

    def window(self, first=0, count=0, chunk=0, max_ctas=0):
        """Restrict the following line passes to the lines [first, first + count) (count == 0: whole slab)."""
        self._chk(self.lib.sgpe_slab_window(self.h, int(first), int(count), int(chunk), int(max_ctas)),
                  'sgpe_slab_window')

    def set_peers(self, peer_ptrs, mode, seg, drow, dplane, base):
        """Destination map of the fused exchange: peer_ptrs[q] = integer address of rank q's buffer."""
        arr = (ctypes.c_void_p * len(peer_ptrs))(*[ctypes.c_void_p(int(v)) for v in peer_ptrs])
        self._chk(self.lib.sgpe_slab_set_peers(self.h, arr, len(peer_ptrs), int(mode), int(seg), int(drow),
                                               int(dplane), int(base)), 'sgpe_slab_set_peers')

    def slab_pack(self, src, dst, lines, nranks, chunk):
        self._chk(self.lib.sgpe_slab_pack(self.h, _dp(src), _dp(dst), int(lines), int(nranks), int(chunk),
                                          self.stream), 'sgpe_slab_pack')

    def slab_unpack(self, src, dst, nranks, block_h, block_w):
        self._chk(self.lib.sgpe_slab_unpack(self.h, _dp(src), _dp(dst), int(nranks), int(block_h), int(block_w),
                                            self.stream), 'sgpe_slab_unpack')

    def accounting(self):
        a, b, c = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_int()
        self._chk(self.lib.sgpe_step_accounting(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)),
                  'sgpe_step_accounting')
        return dict(algorithmic_bytes=a.value, actual_bytes=b.value, launches=c.value)

    def profile_begin(self):
        self._chk(self.lib.sgpe_profile_begin(self.h), 'sgpe_profile_begin')

    def profile_end(self):
        a, b = ctypes.c_double(), ctypes.c_double()
        na, nb = ctypes.c_uint64(), ctypes.c_uint64()
        self._chk(self.lib.sgpe_profile_end(self.h, ctypes.byref(a), ctypes.byref(na), ctypes.byref(b),
                                            ctypes.byref(nb)), 'sgpe_profile_end')
        return dict(col_ms=a.value, col_launches=na.value, row_ms=b.value, row_launches=nb.value)

    def launch_count(self):
        n = ctypes.c_uint64()
        self._chk(self.lib.sgpe_launch_count(self.h, ctypes.byref(n)), 'sgpe_launch_count')
        return n.value
