"""TEST INFRASTRUCTURE: drives tests/emu/libsgpe_emu.so (the kernel sources compiled for the CPU fibre
model) through the *same* C ABI as the CUDA library, with numpy arrays standing in for device memory.
Lets the index math / algebra of the kernels be checked against the oracle on a box without a GPU."""
import ctypes
import os
import subprocess

import numpy as np

from spinor_gpe_b200 import _capi

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def emu_lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, 'emu', 'libsgpe_emu.so')
        subprocess.run(['make', '-s', '-j8', '-C', os.path.join(HERE, 'emu')], check=True)
        _LIB = _capi.bind(ctypes.CDLL(so))
        assert b'emu' in _LIB.sgpe_version()
    return _LIB


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


class EmuPlan:
    """Thin numpy-level wrapper.  Problem arrays follow oracle.Problem (psik (B,2,Ny,Nx) or (2,Ny,Nx))."""

    def __init__(self, nx, ny, batch=1, dtype=np.complex128):
        self.lib = emu_lib()
        self.nx, self.ny, self.batch = nx, ny, batch
        self.cdtype = np.dtype(dtype)
        self.h = ctypes.c_void_p()
        code = _capi.SGPE_C128 if self.cdtype == np.complex128 else _capi.SGPE_C64
        _capi.check(self.lib, self.lib.sgpe_plan_create(ctypes.byref(self.h), nx, ny, batch, code, 0), 'plan_create')
        self.keep = {}

    def close(self):
        if self.h:
            self.lib.sgpe_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _chk(self, rc, what):
        _capi.check(self.lib, rc, what)

    def set_grid(self, dx, dy, dv_r, dv_k, atom_num):
        self._chk(self.lib.sgpe_set_grid(self.h, dx, dy, dv_r, dv_k, atom_num), 'set_grid')

    def set_interactions(self, g):
        self._chk(self.lib.sgpe_set_interactions(self.h, *[float(v) for v in g]), 'set_interactions')

    def _f64(self, key, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        self.keep[key] = a
        return a

    def set_kinetic(self, kin, batched=False):
        a = self._f64('kin', kin)      # (2,Ny,Nx) or (B,2,Ny,Nx)
        plane = self.nx * self.ny
        k0 = ctypes.c_void_p(a.ctypes.data)
        k1 = ctypes.c_void_p(a.ctypes.data + 8 * plane)
        self._chk(self.lib.sgpe_set_kinetic(self.h, k0, k1, 2 * plane if batched else 0), 'set_kinetic')

    def set_potential(self, pot, batched=False, share=False):
        a = self._f64('pot', pot)
        plane = self.nx * self.ny
        p0 = ctypes.c_void_p(a.ctypes.data)
        p1 = p0 if share else ctypes.c_void_p(a.ctypes.data + 8 * plane)
        self._chk(self.lib.sgpe_set_potential(self.h, p0, p1, 2 * plane if batched else 0), 'set_potential')

    def set_kinetic_separable(self, kin_x, kin_y, batched=False):
        kx, ky = self._f64('kin_x', kin_x), self._f64('kin_y', kin_y)
        self._chk(self.lib.sgpe_set_kinetic_separable(self.h, _ptr(kx), _ptr(ky), 2 * self.nx if batched else 0,
                                                      2 * self.ny if batched else 0), 'set_kinetic_separable')

    def set_potential_separable(self, pot_x, pot_y, batched=False):
        px, py = self._f64('pot_x', pot_x), self._f64('pot_y', pot_y)
        self._chk(self.lib.sgpe_set_potential_separable(self.h, _ptr(px), _ptr(py), 2 * self.nx if batched else 0,
                                                        2 * self.ny if batched else 0), 'set_potential_separable')

    def set_coupling(self, mode, coupling=None, omega=None, eiphi=None, batched=False):
        c = self._f64('cpl', coupling) if coupling is not None else None
        o = self._f64('omega', omega) if omega is not None else None
        e = None
        if eiphi is not None:
            e = np.ascontiguousarray(eiphi, dtype=self.cdtype)
            self.keep['eiphi'] = e
        self._chk(self.lib.sgpe_set_coupling(self.h, mode, _ptr(c), self.nx * self.ny if batched else 0,
                                             _ptr(o), _ptr(e)), 'set_coupling')

    def set_option(self, name, value):
        self._chk(self.lib.sgpe_set_option(self.h, name.encode(), int(value)), 'set_option')

    def set_time(self, mode, dt):
        code = _capi.SGPE_TIME_IMAG if mode == 'imag' else _capi.SGPE_TIME_REAL
        self._chk(self.lib.sgpe_set_time(self.h, code, dt), 'set_time')

    def substeps(self):
        a, b = ctypes.c_double(), ctypes.c_double()
        self._chk(self.lib.sgpe_substeps(self.h, ctypes.byref(a), ctypes.byref(b)), 'substeps')
        return a.value, b.value

    def _state(self, arr):
        return np.ascontiguousarray(arr, dtype=self.cdtype).reshape(self.batch, 2, self.ny, self.nx)

    def load(self, psik):
        a = self._state(psik)
        self._chk(self.lib.sgpe_load_psik(self.h, _ptr(a), None), 'load')

    def store(self):
        out = np.empty((self.batch, 2, self.ny, self.nx), dtype=self.cdtype)
        self._chk(self.lib.sgpe_store_psik(self.h, _ptr(out), None), 'store')
        return out

    def single_step(self, dt_sub):
        self._chk(self.lib.sgpe_single_step(self.h, dt_sub, None), 'single_step')

    def full_steps(self, n, want_pops=True):
        pops = np.zeros((self.batch, n, 2)) if want_pops else None
        self._chk(self.lib.sgpe_full_steps(self.h, n, _ptr(pops), 2 * n, 0, None), 'full_steps')
        return pops

    def full_steps_energy(self, n, kl_term=0.0, unwrap=0):
        pops = np.zeros((self.batch, n, 2))
        eng = np.zeros((self.batch, n, 4))
        self._chk(self.lib.sgpe_full_steps_energy(self.h, n, _ptr(pops), 2 * n, 0, _ptr(eng), 4 * n, 0, unwrap, kl_term,
                                                  None), 'full_steps_energy')
        return pops, eng

    def fft2d(self, arr, inverse=False):
        a = self._state(arr)
        out = np.empty_like(a)
        self._chk(self.lib.sgpe_fft2d(self.h, _ptr(a), _ptr(out), int(inverse), None), 'fft2d')
        return out

    def fft1d(self, arr, axis, inverse=False):
        a = self._state(arr)
        out = np.empty_like(a)
        self._chk(self.lib.sgpe_fft1d(self.h, _ptr(a), _ptr(out), axis, int(inverse), None), 'fft1d')
        return out

    def sumsq(self, arr):
        a = self._state(arr)
        out = np.zeros((self.batch, 2))
        self._chk(self.lib.sgpe_sumsq(self.h, _ptr(a), _ptr(out), None), 'sumsq')
        return out

    def normalise(self, arr, vol):
        a = self._state(arr)
        out = np.empty_like(a)
        self._chk(self.lib.sgpe_normalise(self.h, _ptr(a), _ptr(out), vol, None), 'normalise')
        return out

    def unwrap_phase(self, arr, mask=False):
        """arr: (..., Ny, Nx) complex field (kind 0) or float64 wrapped angles (kind 1)."""
        a = np.asarray(arr)
        kind = 0 if np.iscomplexobj(a) else 1
        a = np.ascontiguousarray(a, dtype=self.cdtype if kind == 0 else np.float64)
        nplanes = a.size // (self.nx * self.ny)
        out = np.empty(a.shape, dtype=np.float64)
        self._chk(self.lib.sgpe_unwrap_phase(self.h, _ptr(a), kind, nplanes, int(mask), _ptr(out), None), 'unwrap_phase')
        return out

    def kinetic_spectral(self, psik=None):
        a = self._state(psik) if psik is not None else None
        out = np.zeros((self.batch, 2))
        self._chk(self.lib.sgpe_kinetic_spectral(self.h, _ptr(a), _ptr(out), None), 'kinetic_spectral')
        return out

    def gradient(self, field, h0, h1):
        is_c = np.iscomplexobj(field)
        rd = np.float64 if self.cdtype == np.complex128 else np.float32
        f = np.ascontiguousarray(field, dtype=self.cdtype if is_c else rd)
        g0, g1 = np.empty_like(f), np.empty_like(f)
        self._chk(self.lib.sgpe_gradient(self.h, _ptr(f), int(is_c), float(h0), float(h1), _ptr(g0), _ptr(g1), None), 'gradient')
        return [g0, g1]

    def energy(self, psik=None, kl_term=0.0, unwrap=0):
        a = self._state(psik) if psik is not None else None
        out = np.zeros((self.batch, 4))
        self._chk(self.lib.sgpe_energy(self.h, _ptr(a), unwrap, kl_term, _ptr(out), None), 'energy')
        return out

    def run_host(self, psik, n):
        a = self._state(psik)
        out = np.empty_like(a)
        pops = np.zeros((self.batch, n, 2))
        self._chk(self.lib.sgpe_run_host(self.h, _ptr(a), _ptr(out), n, _ptr(pops), None), 'run_host')
        return out, pops


from spinor_gpe_b200._separable import split_separable  # noqa: E402


def plan_from_problem(prob, mode, dt, dtype=np.complex128, separable=False):
    """Configure an EmuPlan from an oracle.Problem the way the product host code configures a CUDA plan
    (dense operators; uniform coupling detected; exp(i*expon) along x)."""
    ny, nx = prob.psik.shape[-2:]
    pl = EmuPlan(nx, ny, 1, dtype)
    pl.set_grid(prob.dr[0], prob.dr[1], prob.dv_r, prob.dv_k, prob.atom_num)
    pl.set_interactions((prob.g_uu, prob.g_dd, prob.g_ud))
    pot = prob.pot.numpy()
    ksep = split_separable(prob.kin.numpy()) if separable else None
    psep = split_separable(pot) if separable else None
    if ksep is not None:
        pl.set_kinetic_separable(*ksep)
    else:
        pl.set_kinetic(prob.kin.numpy())
    pl.set_potential(pot, share=bool(np.array_equal(pot[0], pot[1])))
    if psep is not None:
        pl.set_potential_separable(*psep)
    pl.sep_used = (ksep is not None, psep is not None)
    if prob.is_coupling:
        cpl = prob.coupling.numpy()
        eiphi = None
        if prob.expon.ndim == 2:
            eiphi = np.exp(1j * prob.expon.numpy()[0])
        if np.all(cpl == cpl.flat[0]):
            pl.set_coupling(_capi.SGPE_COUPLING_UNIFORM, omega=np.array([cpl.flat[0]]), eiphi=eiphi)
        else:
            pl.set_coupling(_capi.SGPE_COUPLING_DENSE, coupling=cpl, eiphi=eiphi)
    else:
        pl.set_coupling(_capi.SGPE_COUPLING_NONE)
    pl.set_time(mode, dt)
    pl.load(prob.psik.numpy())
    return pl


class EmuTorchPlan:
    """The subset of spinor_gpe_b200.plan.Plan that sweep.run_sweep uses, on top of the emulated library
    with CPU tensors / numpy arrays (lets the multi-process host logic run under gloo without a GPU)."""

    def __init__(self, nx, ny, batch=1, dtype=None, device='cpu'):
        import torch
        cd = np.complex64 if dtype == torch.complex64 else np.complex128
        self.p = EmuPlan(nx, ny, batch, cd)
        self.nx, self.ny, self.batch = nx, ny, batch

    def set_grid(self, *a):
        self.p.set_grid(*[float(v) for v in a])

    def set_interactions(self, *g):
        self.p.set_interactions(g)

    def set_kinetic(self, k0, k1, batched=False):
        self.p.set_kinetic(np.stack([np.asarray(k0), np.asarray(k1)]), batched=False)

    def set_kinetic_separable(self, kx, ky, batched=False):
        self.p.set_kinetic_separable(kx, ky, batched)

    def set_potential(self, p0, p1, batched=False, shared=False):
        p0, p1 = np.asarray(p0), np.asarray(p1)
        self.p.set_potential(np.stack([p0, p1], axis=1 if batched else 0), batched=batched)

    def set_potential_separable(self, px, py, batched=False):
        self.p.set_potential_separable(px, py, batched)

    def set_coupling(self, mode, coupling=None, omega=None, eiphi=None, batched=False):
        self.p.set_coupling(mode, coupling=coupling, omega=omega, eiphi=eiphi, batched=batched)

    def set_time(self, mode, dt):
        self.p.set_time(mode, dt)

    def load(self, psik):
        self.p.load(psik)

    def full_steps(self, n, pops=None, first=0):
        arr = pops.numpy() if pops is not None else None
        stride = arr.shape[1] * 2 if arr is not None else 0
        self.p._chk(self.p.lib.sgpe_full_steps(self.p.h, n, _ptr(arr), stride, first, None), 'full_steps')

    def energy(self, psik=None, kl_term=0.0, unwrap='none'):
        import torch
        from spinor_gpe_b200.plan import UNWRAP_MODES
        return torch.from_numpy(self.p.energy(psik, kl_term, UNWRAP_MODES[unwrap]))

    def store(self):
        import torch
        return torch.from_numpy(self.p.store())

    def close(self):
        self.p.close()
