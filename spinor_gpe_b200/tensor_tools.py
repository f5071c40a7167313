"""Drop-in for the reference's ``spinor_gpe/pspinor/tensor_tools.py`` on the propagator path.

Wavefunctions are, as in the reference, lists of two (Ny, Nx) arrays.  Two kinds of input exist:

* **CUDA tensors** (the propagator path) — every operation runs in the hand-written kernels of
  ``libsgpe.so`` through the C ABI.  CPU tensors are rejected: there is no torch/CPU fallback.
* **NumPy arrays** — host-side problem set-up and result analysis (``PSpinor``, ``PropResult``), which the
  reference also does in NumPy (its ``isinstance(..., np.ndarray)`` branches).  These never touch the
  time-stepping path.
"""
import operator
from collections import OrderedDict
from functools import reduce

import numpy as np
import torch

from . import _capi
from .plan import Plan

# Helper plans of the stand-alone transforms / reductions: twiddles and reduction scratch only (a plan allocates its
# working state on the first sgpe_load_psik, which these never call), a handful kept, least recently used evicted.
_PLANS = OrderedDict()
_MAX_PLANS = 8


def _cached_plan(nx, ny, dtype, device):
    from .plan import supported_mesh
    for name, n in (('x', nx), ('y', ny)):
        ok, why = supported_mesh(n)
        if ok and n > 4096:
            ok, why = False, "the stand-alone helpers transform lines of up to 4096 points"
        if not ok:
            raise ValueError(f"{n} mesh points along {name}: {why}")
    key = (nx, ny, dtype, device.index)
    pl = _PLANS.pop(key, None)
    if pl is None:
        pl = Plan(nx, ny, 1, dtype, device)
        while len(_PLANS) >= _MAX_PLANS:
            _PLANS.popitem(last=False)[1].close()
    _PLANS[key] = pl
    return pl


def _plan_for(t):
    """A cached bare plan (transforms / reductions only) matching a CUDA tensor's shape and dtype."""
    ny, nx = t.shape[-2:]
    return _cached_plan(nx, ny, t.dtype, t.device)


def _is_np(psi):
    first = psi[0] if isinstance(psi, (list, tuple)) else psi
    return isinstance(first, np.ndarray)


def _stack_cuda(psi):
    if not isinstance(psi, (list, tuple)) or len(psi) != 2:
        raise TypeError("a wavefunction is a list of two (Ny, Nx) arrays")
    for p in psi:
        if not isinstance(p, torch.Tensor):
            raise TypeError(f"expected torch tensors, got {type(p)}")
        if not p.is_cuda:
            raise RuntimeError("spinor_gpe_b200 computes on CUDA tensors only (no CPU fallback); "
                               "move the wavefunction to the GPU or pass NumPy arrays for host analysis")
    dt = torch.complex64 if psi[0].dtype in (torch.complex64, torch.float32) else torch.complex128
    return torch.stack([p.to(dt) for p in psi]).contiguous()


def _unstack(t):
    t = t.reshape(2, t.shape[-2], t.shape[-1])
    return [t[0], t[1]]


# ----------------------------------------------------------------------------- conversions
def to_numpy(input_tens):
    """tensor_tools.py:14-38."""
    if isinstance(input_tens, list):
        return [inp.cpu().numpy() for inp in input_tens]
    return input_tens.cpu().numpy()


def to_tensor(input_arr, dev='cuda', dtype=64):
    """tensor_tools.py:41-77 — dtype code 32 / 64 / 128."""
    kinds = {32: torch.float32, 64: torch.float64, 128: torch.complex128}
    if isinstance(input_arr, list):
        return [torch.as_tensor(a, dtype=kinds[dtype], device=dev) for a in input_arr]
    return torch.as_tensor(input_arr, dtype=kinds[dtype], device=dev)


def to_cpu(input_tens):
    """tensor_tools.py:80-100."""
    if isinstance(input_tens, list):
        return [t.cpu() for t in input_tens]
    return input_tens.cpu()


def to_gpu(input_tens, dev='cuda'):
    """tensor_tools.py:103-127."""
    if isinstance(input_tens, list):
        return [t.to(dev) for t in input_tens]
    return input_tens.to(dev)


def prod(factors):
    """tensor_tools.py:594-600."""
    return reduce(operator.mul, factors, 1)


# ----------------------------------------------------------------------------- transforms
def _dr_pair(delta_r):
    if isinstance(delta_r, torch.Tensor):
        delta_r = delta_r.detach().cpu().numpy()
    return float(delta_r[0]), float(delta_r[1])


def _gpu_transform(psi, delta_r, kind, axis=None):
    t = _stack_cuda(psi)
    pl = _plan_for(t)
    dx, dy = _dr_pair(delta_r)
    pl.set_grid(dx, dy, dx * dy, 1.0, 1.0)
    if kind in ('fft2', 'ifft2'):
        out = pl.fft2d(t, inverse=(kind == 'ifft2'))
    else:
        out = pl.fft1d(t, axis, inverse=(kind == 'ifft1'))
    return _unstack(out)


def fft_2d(psi, delta_r=(1, 1)):
    """tensor_tools.py:201-228: fftn x (dx dy / 2 pi), fftshift."""
    if _is_np(psi):
        s = prod(delta_r) / (2 * np.pi)
        return [np.fft.fftshift(np.fft.fftn(p) * s) for p in psi]
    return _gpu_transform(psi, delta_r, 'fft2')


def ifft_2d(psik, delta_r=(1, 1)):
    """tensor_tools.py:231-258."""
    if _is_np(psik):
        s = prod(delta_r) / (2 * np.pi)
        return [np.fft.ifftn(np.fft.ifftshift(p)) / s for p in psik]
    return _gpu_transform(psik, delta_r, 'ifft2')


def fft_1d(psi, delta_r=(1, 1), axis=0):
    """tensor_tools.py:130-164 (axis 0 -> x, the last array dimension)."""
    if _is_np(psi):
        ax = 1 - axis
        s = delta_r[axis] / np.sqrt(2 * np.pi)
        return [np.fft.fftshift(np.fft.fft(p, axis=ax) * s, axes=ax) for p in psi]
    return _gpu_transform(psi, delta_r, 'fft1', axis)


def ifft_1d(psik, delta_r=(1, 1), axis=0):
    """tensor_tools.py:167-198."""
    if _is_np(psik):
        ax = 1 - axis
        s = delta_r[axis] / np.sqrt(2 * np.pi)
        return [np.fft.ifft(np.fft.ifftshift(p, axes=ax), axis=ax) / s for p in psik]
    return _gpu_transform(psik, delta_r, 'ifft1', axis)


# ----------------------------------------------------------------------------- densities / reductions
def norm_sq(psi_comp):
    """tensor_tools.py:414-441."""
    if isinstance(psi_comp, np.ndarray):
        return np.abs(psi_comp) ** 2
    if isinstance(psi_comp, torch.Tensor):
        if not psi_comp.is_cuda:
            raise RuntimeError("CUDA tensors only (no CPU fallback)")
        v = torch.view_as_real(psi_comp) if psi_comp.is_complex() else psi_comp.unsqueeze(-1)
        return (v * v).sum(-1)
    raise TypeError(f"`psi_comp` is of type {type(psi_comp)}")


def density(psi):
    """tensor_tools.py:392-411."""
    if isinstance(psi, list):
        return [norm_sq(p) for p in psi]
    return norm_sq(psi)


def calc_pops(psi, vol_elem=1.0):
    """tensor_tools.py:466-484 — per-component atom numbers (fused sum-of-squares kernel on the GPU)."""
    if _is_np(psi):
        return [float((np.abs(p) ** 2).sum() * vol_elem) for p in psi]
    t = _stack_cuda(psi)
    sums = _plan_for(t).sumsq(t)[0].cpu().numpy()
    return [float(sums[0] * float(vol_elem)), float(sums[1] * float(vol_elem))]


def calc_atoms(psi, vol_elem=1.0):
    """tensor_tools.py:444-463."""
    return sum(calc_pops(psi, vol_elem))


def norm(psi, vol_elem, atom_num, pop_frac=None):
    """tensor_tools.py:261-311 — returns (psi / sqrt(nf), dens / nf)."""
    if pop_frac is not None:
        raise NotImplementedError("Normalizing to the expected population fractions is not implemented "
                                  "(as in the reference, tensor_tools.py:296-309).")
    if _is_np(psi):
        dens = [np.abs(p) ** 2 for p in psi]
        nf = np.sum(dens[0] + dens[1]) * vol_elem / atom_num
        return [p / np.sqrt(nf) for p in psi], [d / nf for d in dens]
    t = _stack_cuda(psi)
    pl = _plan_for(t)
    pl.set_grid(1.0, 1.0, 1.0, 1.0, float(atom_num))
    out = _unstack(pl.normalise(t, float(vol_elem)))
    return out, [norm_sq(p) for p in out]


# ----------------------------------------------------------------------------- host-side analysis (NumPy, as in the reference)
def grad_comp(psi_comp, delta_r):
    """tensor_tools.py:331-350.  NumPy arrays go to np.gradient as in the reference.  Beyond the reference (which raises
    for tensors, :343-345), a CUDA tensor is differentiated on the device (``sgpe_gradient``: the same stencils) and
    the two derivatives stay there."""
    if isinstance(psi_comp, np.ndarray):
        return np.gradient(psi_comp, *np.array(delta_r))
    if isinstance(psi_comp, torch.Tensor):
        if not psi_comp.is_cuda:
            raise RuntimeError("spinor_gpe_b200 computes on CUDA tensors only (no CPU fallback)")
        if psi_comp.dim() != 2:
            raise ValueError("grad_comp takes one (Ny, Nx) component")
        h0, h1 = (float(d) for d in np.array(delta_r))
        cdtype = torch.complex64 if psi_comp.dtype in (torch.complex64, torch.float32) else torch.complex128
        ny, nx = psi_comp.shape
        return _cached_plan(nx, ny, cdtype, psi_comp.device).gradient(psi_comp, h0, h1)
    raise TypeError(f"`psi_comp` is of type {type(psi_comp)}")


def grad(psi, delta_r):
    """tensor_tools.py:314-328."""
    if isinstance(psi, list):
        return [grad_comp(p, delta_r) for p in psi]
    return grad_comp(psi, delta_r)


def grad_sq_comp(psi_comp, delta_r):
    """tensor_tools.py:353-357."""
    g0, g1 = grad_comp(psi_comp, delta_r)
    return g0 ** 2 + g1 ** 2


def grad_sq(psi, delta_r):
    """tensor_tools.py:360-367."""
    if isinstance(psi, list):
        return [grad_sq_comp(p, delta_r) for p in psi]
    return grad_sq_comp(psi, delta_r)


def conj_comp(psi_comp):
    """tensor_tools.py:379-389."""
    if isinstance(psi_comp, np.ndarray):
        return np.conj(psi_comp)
    if isinstance(psi_comp, torch.Tensor):
        return torch.conj(psi_comp)
    raise TypeError(f"`psi_comp` is of type {type(psi_comp)}")


def conj(psi):
    """tensor_tools.py:370-376."""
    if isinstance(psi, list):
        return [conj_comp(p) for p in psi]
    return conj_comp(psi)


def _unwrap_plan(ny, nx):
    """A cached bare complex128 plan on the current CUDA device for the (ny, nx) mesh (raises without CUDA)."""
    from ._lib import require_cuda
    require_cuda()
    return _cached_plan(nx, ny, torch.complex128, torch.device('cuda', torch.cuda.current_device()))


def _unwrap_2d(ang):
    """The reference calls skimage.restoration.unwrap_phase (tensor_tools.py:531); here the same algorithm
    (Herraez et al. 2002, reliability-sorted region merging) runs through ``sgpe_unwrap_phase``: the per-pixel work
    and the edge sort on the GPU, the region merging in the library's host code.  NumPy in, NumPy out."""
    ang = np.ascontiguousarray(ang, dtype=np.float64)
    out = _unwrap_plan(*ang.shape).unwrap_phase(torch.from_numpy(ang))
    return out.cpu().numpy()


def phase_comp(psi_comp, uwrap=False, dens=None):
    """tensor_tools.py:514-539.  Beyond the reference (which raises for tensors, :533-536), a CUDA tensor can be
    unwrapped too; the result stays on the device."""
    if isinstance(psi_comp, np.ndarray):
        ang = np.angle(psi_comp)
        if uwrap:
            ang = _unwrap_2d(ang)
    elif isinstance(psi_comp, torch.Tensor):
        if uwrap:
            if not psi_comp.is_cuda:
                raise RuntimeError("spinor_gpe_b200 computes on CUDA tensors only (no CPU fallback)")
            ang = _plan_for(psi_comp.to(torch.complex128)).unwrap_phase(psi_comp.to(torch.complex128))
        else:
            ang = torch.angle(psi_comp)
    else:
        raise TypeError(f"`psi_comp` is of type {type(psi_comp)}")
    if dens is not None:
        ang[dens < (dens.max() * 1e-6)] = 0
    return ang


def phase(psi, uwrap=False, dens=None):
    """tensor_tools.py:487-511."""
    if isinstance(psi, list):
        if dens is None:
            dens = [None] * len(psi)
        assert len(psi) == len(dens), "`psi` and `dens` should have the same length."
        return [phase_comp(p, uwrap, d) for p, d in zip(psi, dens)]
    return phase_comp(psi, uwrap, dens)


# ----------------------------------------------------------------------------- operator tables
# The fused kernels evaluate the evolution operators in registers; these two build the explicit
# tables only when a caller asks for them (TensorPropagator.eng_out / eng_in views, tests).
def evolution_op(t_step, energy):
    """tensor_tools.py:546-560 — exp(-i E t) (element-wise torch ops on whatever device E lives on)."""
    if isinstance(energy, list):
        return [torch.exp(-1.0j * e * t_step) for e in energy]
    return torch.exp(-1.0j * energy * t_step)


def coupling_op(t_step, coupling=None, expon=None):
    """tensor_tools.py:563-591."""
    if coupling is None:
        coupling = torch.tensor(0.0)
    if expon is None:
        expon = torch.tensor(0.0, device=coupling.device)
    arg = coupling * t_step / 2
    cosine = torch.cos(arg)
    sine = -1.0j * torch.sin(arg)
    return [[cosine, sine * torch.exp(-1.0j * expon)], [sine * torch.exp(1.0j * expon), cosine]]


def inner_prod():
    """tensor_tools.py:542 (a stub in the reference as well)."""


def expect_val(psi):
    """tensor_tools.py:603-607."""
    raise NotImplementedError("Function for computing the expectation value of an arbitrary spatial "
                              "operator is not implemented.")
