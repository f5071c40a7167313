"""Drop-in for the reference's ``TensorPropagator`` (spinor_gpe/pspinor/tensor_propagator.py).

Same constructor, methods and attributes; the time stepping runs in the fused sm_100a kernels of
``libsgpe.so``.  There is no CPU path: ``device`` must be a CUDA device.

Differences a user can observe (all documented in DESIGN.md):

* ``psik`` is materialised lazily — between ``full_step()`` calls the state lives in the kernels'
  internal (k_x, y) layout and un-normalised; reading ``prop.psik`` closes the junction and returns
  the normalised k-space state the reference would hold (tensor_propagator.py:271).
* ``eng_out`` / ``eng_in`` are views that build the explicit operator tables (tensor_propagator.py:138-149)
  only when indexed; the kernels evaluate the operators in registers.
* when the kinetic / potential grids are separable (``g[y, x] = gx[x] + gy[y]``, true for everything
  ``PSpinor`` builds by default) the kernels use 1-D factor tables instead of per-point exp / sincos;
  ``separable=False`` forces the general dense path.
* ``eng_expect`` runs on the GPU.  The reference unwraps the phase of each component with
  ``skimage.restoration.unwrap_phase`` (tensor_tools.py:531); ``unwrap='herraez'`` is that algorithm
  (``sgpe_unwrap_phase``: per-pixel work, the edge sort and the region merging on the device), ``'none'`` differentiates the wrapped phase as it is and ``'local'`` uses locally wrapped
  differences (both stay on the device and never synchronise).  The constructor's ``unwrap`` argument
  (default ``DEFAULT_UNWRAP``) is what ``prop_loop`` uses for ``PropResult.eng_final``.
"""
import os

import numpy as np
import torch

from . import _capi
from . import tensor_tools as ttools
from .plan import Plan
from .plotting_tools import next_available_path      # names the sampled-wavefunction file (:199-200)
from .prop_result import PropResult
from ._separable import split_separable

try:                                        # progress bar as in the reference (tensor_propagator.py:185)
    from tqdm import tqdm as _tqdm
except ImportError:                         # pragma: no cover
    _tqdm = None

MAGIC_GAMMA = 1 / (2 + 2 ** (1 / 3))        # tensor_propagator.py:101
MAX_LINE = 4096                             # longest line one CTA transforms (sgpe_plan_create)
# phase treatment of eng_expect when the caller does not choose one (see the module docstring)
DEFAULT_UNWRAP = 'herraez'


def check_mesh(nx, ny):
    """Mesh sizes the kernels transform: the reference asserts even sizes (pspinor.py:331-332); see ``supported_mesh``."""
    from .plan import supported_mesh
    for name, n in (('x', nx), ('y', ny)):
        ok, why = supported_mesh(n)
        if not ok:
            raise ValueError(f"mesh_points along {name} = {n}: {why}")


def _grid(arr, ny, nx, name):
    """C-contiguous float64 (Ny, Nx) copy / view of a user-supplied operator grid."""
    a = np.ascontiguousarray(np.asarray(arr), dtype=np.float64)
    if a.shape != (ny, nx):
        raise ValueError(f"{name} has shape {a.shape}, the mesh is ({ny}, {nx})")
    return a


def _to_host(comps):
    """Two (Ny, Nx) device tensors -> two NumPy arrays, through page-locked staging (the pageable path of
    ``Tensor.cpu()`` runs at ~2 GB/s; torch keeps the pinned blocks for the next call)."""
    host = [torch.empty(c.shape, dtype=c.dtype, pin_memory=True) for c in comps]
    for h, c in zip(host, comps):
        h.copy_(c, non_blocking=True)
    torch.cuda.current_stream(comps[0].device).synchronize()
    return [h.numpy() for h in host]


class _LazyTensors(dict):
    """dict whose values are uploaded to the GPU on first access."""

    def __init__(self, source, keys, device):
        super().__init__()
        self._source, self._keys, self._device = source, tuple(keys), device

    def __missing__(self, key):
        if key not in self._keys:
            raise KeyError(key)
        val = torch.tensor(np.asarray(self._source[key]), device=self._device)
        self[key] = val
        return val

    def keys(self):
        return list(self._keys)

    def __contains__(self, key):
        return key in self._keys


class _OperatorView:
    """``eng_out`` / ``eng_in``: behaves like the reference's dict {'kin', 'pot', 'coupl'} of explicit
    operator tables, built on demand (never used by the fused kernels)."""

    def __init__(self, prop, dt_sub, outer):
        self._prop, self.dt_sub, self.outer = prop, dt_sub, outer
        self._cache = {}

    def __getitem__(self, key):
        if key not in self._cache:
            p = self._prop
            if key == 'kin':
                self._cache[key] = ttools.evolution_op(self.dt_sub / 2, p.kin_eng_spin)
            elif key == 'pot':
                self._cache[key] = ttools.evolution_op(self.dt_sub, p.pot_eng_spin)
            elif key == 'coupl':
                if self.outer:      # tensor_propagator.py:142-143
                    self._cache[key] = ttools.coupling_op(self.dt_sub, p.coupling / 2, p.expon)
                else:               # tensor_propagator.py:148-149
                    self._cache[key] = ttools.coupling_op(self.dt_sub / 2, p.coupling, p.expon)
            else:
                raise KeyError(key)
        return self._cache[key]

    def keys(self):
        return ['kin', 'pot', 'coupl']


class TensorPropagator:
    """GPU propagator of the pseudospin-1/2 GPE (see the module docstring; attributes as in the
    reference, tensor_propagator.py:14-59)."""

    # pylint: disable=too-many-instance-attributes
    def __init__(self, spin, t_step, n_steps, device='cuda', time='imag', is_sampling=False, n_samples=1,
                 precision='c128', progress=False, separable='auto', unwrap=None, long_lines=None,
                 track_energy=False):
        dev = torch.device(device)
        # track_energy: prop_loop also records eng_expect of every step (PropResult.eng_history, (n_steps, 4)); pass
        # 'none', 'local' or 'herraez' to choose the phase treatment (True = 'none').  'none' / 'local' are fused with
        # the stepping and never synchronise; 'herraez' (the reference's definition) unwraps the phase after every step
        # (device-side region merging steered from the host: tens of milliseconds per step at 2048^2)
        self.track_energy = 'none' if track_energy is True else (track_energy or None)
        if self.track_energy not in (None, 'none', 'local', 'herraez'):
            raise ValueError("track_energy must be False, True, 'none', 'local' or 'herraez'")
        self.unwrap = DEFAULT_UNWRAP if unwrap is None else unwrap
        if self.unwrap not in ('none', 'local', 'herraez'):
            raise ValueError("unwrap must be 'none', 'local' or 'herraez'")
        if dev.type != 'cuda':
            raise RuntimeError(f"device={device!r}: the B200 propagator has no CPU path; pass a CUDA device")
        if dev.index is None:
            dev = torch.device('cuda', torch.cuda.current_device())
        self.n_steps = n_steps
        self.device = device
        self._dev = dev
        self.paths = spin.paths
        self._progress = progress
        self._separable_opt = bool(separable)

        if time == 'imag':                                  # tensor_propagator.py:96-99
            self.t_step = -1.0j * t_step
        elif time == 'real':
            self.t_step = t_step
        else:
            raise ValueError("time must be 'imag' or 'real'")
        self._time = time
        self._dt = float(t_step)
        self.dt_out = self.t_step * MAGIC_GAMMA             # :102
        self.dt_in = self.t_step * (1 - 2 * MAGIC_GAMMA)    # :103

        self.rand_seed = spin.rand_seed
        if self.rand_seed is not None:
            torch.manual_seed(self.rand_seed)               # :105-107 (no RNG is consumed afterwards)
        self.is_sampling = is_sampling

        self.atom_num = spin.atom_num
        self.is_coupling = spin.is_coupling
        self.g_sc = spin.g_sc
        self.kL_recoil = spin.kL_recoil                     # pylint: disable=invalid-name
        self._rot_coupling = bool(spin.rot_coupling)
        cdtype = {'c128': torch.complex128, 'c64': torch.complex64}[precision]
        self._cdtype = cdtype

        psik0 = [np.asarray(spin.psik[0]), np.asarray(spin.psik[1])]      # (no stacked host copy: 134 MB at 2048^2)
        ny, nx = psik0[0].shape[-2:]
        if psik0[1].shape != psik0[0].shape or psik0[0].ndim != 2:
            raise ValueError("psik must be two (Ny, Nx) arrays")
        check_mesh(nx, ny)
        # the kernels read row-major (Ny, Nx) float64 grids through raw pointers: whatever layout the user's arrays
        # have (Fortran order, transposed views, other dtypes), the device copies are C-contiguous float64
        # (uploaded when first needed: with separable grids the kernels work from 1-D factor tables and the 2-D device
        # copies exist only if somebody reads the public attributes kin_eng_spin / pot_eng_spin / coupling)
        self._kin_np = [_grid(k, ny, nx, 'kin_eng_spin') for k in spin.kin_eng_spin]
        pot_shared = spin.pot_eng_spin[0] is spin.pot_eng_spin[1] or np.array_equal(spin.pot_eng_spin[0],
                                                                                    spin.pot_eng_spin[1])
        p0 = _grid(spin.pot_eng_spin[0], ny, nx, 'pot_eng_spin')
        self._pot_np = [p0, p0] if pot_shared else [p0, _grid(spin.pot_eng_spin[1], ny, nx, 'pot_eng_spin')]
        self._kin_dev = self._pot_dev = self._cpl_dev = None
        keys_space = ['dr', 'dk', 'x_mesh', 'y_mesh', 'dv_r', 'dv_k']
        self.space = _LazyTensors(spin.space, keys_space, dev)
        self._dr = (float(spin.space['dr'][0]), float(spin.space['dr'][1]))
        self._dv_r, self._dv_k = float(spin.space['dv_r']), float(spin.space['dv_k'])
        cpl_np = self._cpl_np = _grid(spin.coupling, ny, nx, 'coupling')

        if self.is_sampling:                                # :131-134
            assert self.n_steps % n_samples == 0, (
                f"The number of samples requested {n_samples} does not evenly "
                f"divide the total number of steps {self.n_steps}.")
        self.sample_rate = self.n_steps / n_samples         # :136

        # ---- the plan
        # meshes with more than 4096 points along a line do not fit one CTA per line: they run on the long-line
        # machinery of the slab mode (four-step lines) on this one device.  ``long_lines`` = dict of SlabPropagator
        # keywords forces that path on a small mesh (tests: split_x / split_y).
        self._long = long_lines is not None or max(nx, ny) > MAX_LINE
        if self._long:
            from .slab import LongLinePlan
            self._plan = LongLinePlan(spin, self._dt, time, dev, precision, **(long_lines or {}))
            self.separable = {'kin': True, 'pot': None}
        else:
            self._plan = pl = Plan(nx, ny, 1, cdtype, dev)
            pl.set_grid(self._dr[0], self._dr[1], self._dv_r, self._dv_k, self.atom_num)
            pl.set_interactions(self.g_sc['uu'], self.g_sc['dd'], self.g_sc['ud'])
            self._bind_operators(cpl_np, spin)
            pl.set_time(time, self._dt)
        self._psik_cache = None
        self.psik = ttools.to_tensor([psik0[0], psik0[1]], dev=dev, dtype=128)

        self.eng_out = _OperatorView(self, self.dt_out, outer=True)      # :138-143
        self.eng_in = _OperatorView(self, self.dt_in, outer=False)       # :144-149

    # ------------------------------------------------------------------ public 2-D operator grids (tensor_propagator.py:111-122)
    @property
    def kin_eng_spin(self):
        if self._kin_dev is None:
            self._kin_dev = ttools.to_tensor(self._kin_np, dev=self._dev)
        return self._kin_dev

    @kin_eng_spin.setter
    def kin_eng_spin(self, value):
        self._kin_dev = value

    @property
    def pot_eng_spin(self):
        if self._pot_dev is None:
            if self._pot_np[0] is self._pot_np[1]:
                p0 = ttools.to_tensor(self._pot_np[0], dev=self._dev)
                self._pot_dev = [p0, p0]
            else:
                self._pot_dev = ttools.to_tensor(self._pot_np, dev=self._dev)
        return self._pot_dev

    @pot_eng_spin.setter
    def pot_eng_spin(self, value):
        self._pot_dev = value

    @property
    def coupling(self):
        if self._cpl_dev is None:
            self._cpl_dev = ttools.to_tensor(self._cpl_np, dev=self._dev)
        return self._cpl_dev

    @coupling.setter
    def coupling(self, value):
        self._cpl_dev = value

    # ------------------------------------------------------------------ set-up helpers
    def _bind_operators(self, cpl_np, spin):
        import ctypes
        pl = self._plan
        # separable fast path (1-D factor tables) when the grids allow it; the dense path stays general
        self.separable = {'kin': False, 'pot': False}
        ksep = psep = None
        if self._separable_opt:
            ksep = split_separable(self._kin_np)
            if self._pot_np[0] is self._pot_np[1]:              # one grid for both components: checked once
                psep = split_separable(self._pot_np[:1])
                if psep is not None:
                    psep = (np.repeat(psep[0], 2, axis=0), np.repeat(psep[1], 2, axis=0))
            else:
                psep = split_separable(self._pot_np)
        # the plan borrows the device pointers of the dense grids: it keeps the tensors alive itself, so rebinding
        # the public attributes (prop.pot_eng_spin = ...) cannot leave it with a dangling pointer
        if ksep is not None:
            pl.set_kinetic_separable(*ksep)
            self.separable['kin'] = True
        else:
            k0, k1 = self.kin_eng_spin
            assert k0.is_contiguous() and k1.is_contiguous() and k0.dtype == torch.float64
            pl.keep['kin_ref'] = (k0, k1)
            pl._chk(pl.lib.sgpe_set_kinetic(pl.h, ctypes.c_void_p(k0.data_ptr()), ctypes.c_void_p(k1.data_ptr()), 0),
                    'sgpe_set_kinetic')
        if psep is not None:
            pl.set_potential_separable(*psep)
            self.separable['pot'] = True
        else:
            p0, p1 = self.pot_eng_spin
            assert p0.is_contiguous() and p1.is_contiguous() and p0.dtype == torch.float64
            pl.keep['pot_ref'] = (p0, p1)
            pl._chk(pl.lib.sgpe_set_potential(pl.h, ctypes.c_void_p(p0.data_ptr()), ctypes.c_void_p(p1.data_ptr()), 0),
                    'sgpe_set_potential')
        eiphi = None
        if self.is_coupling and not self._rot_coupling:
            # expon = 2 kL x_mesh (tensor_propagator.py:129) depends on x only: ship exp(i expon) along x
            x_mesh = np.asarray(spin.space['x_mesh'])
            if not np.array_equal(x_mesh, np.broadcast_to(x_mesh[0], x_mesh.shape)):
                raise NotImplementedError("x_mesh must be constant along y (a meshgrid of space['x'])")
            eiphi = np.exp(1j * 2 * self.kL_recoil * x_mesh[0])
        self._eiphi = eiphi
        if not self.is_coupling or not np.any(cpl_np):
            # Omega == 0 everywhere: the coupling operator is exactly the identity (cos 0 = 1, sin 0 = 0)
            pl.set_coupling(_capi.SGPE_COUPLING_NONE)
            if np.any(cpl_np):
                # is_coupling False with a non-zero grid: the step skips the operator (tensor_propagator.py:252, 258)
                # but eng_expect still adds the coupling energy from self.coupling (:319-321)
                if np.all(cpl_np == cpl_np.flat[0]):
                    pl.set_energy_coupling(_capi.SGPE_COUPLING_UNIFORM, omega=np.array([cpl_np.flat[0]]))
                else:
                    pl.set_energy_coupling(_capi.SGPE_COUPLING_DENSE, coupling=self.coupling)
        elif np.all(cpl_np == cpl_np.flat[0]):
            pl.set_coupling(_capi.SGPE_COUPLING_UNIFORM, omega=np.array([cpl_np.flat[0]]), eiphi=eiphi)
        else:
            pl.keep['coupling_ref'] = self.coupling
            pl._chk(pl.lib.sgpe_set_coupling(pl.h, _capi.SGPE_COUPLING_DENSE,
                                             ctypes.c_void_p(self.coupling.data_ptr()), 0, None,
                                             ctypes.c_void_p(self._upload_eiphi(eiphi))), 'sgpe_set_coupling')

    def _upload_eiphi(self, eiphi):
        if eiphi is None:
            return None
        t = torch.as_tensor(eiphi).to(device=self._dev, dtype=self._cdtype).contiguous()
        self._plan.keep['eiphi'] = t
        return t.data_ptr()

    @property
    def expon(self):
        """tensor_propagator.py:126-129."""
        if self._rot_coupling:
            return torch.tensor(0.0, device=self._dev)
        return 2 * self.kL_recoil * self.space['x_mesh']

    # ------------------------------------------------------------------ state
    @property
    def psik(self):
        """The normalised k-space wavefunction (list of two (Ny, Nx) CUDA tensors)."""
        if self._psik_cache is None:
            out = self._plan.store()
            if out.dtype != torch.complex128:
                out = out.to(torch.complex128)
            self._psik_cache = [out[0, 0], out[0, 1]]
        return self._psik_cache

    @psik.setter
    def psik(self, value):
        if isinstance(value, (list, tuple)):
            comps = [torch.as_tensor(v) if not isinstance(v, torch.Tensor) else v for v in value]
            value = torch.stack([c.to(self._dev) for c in comps])
        self._plan.load(value.reshape(1, 2, *value.shape[-2:]))
        self._psik_cache = None

    # ------------------------------------------------------------------ stepping
    def full_step(self):
        """tensor_propagator.py:214-222 — three single steps with the magic-gamma sub-steps."""
        self._plan.full_steps(1)
        self._psik_cache = None

    def single_step(self, t_step, eng=None):
        """tensor_propagator.py:224-271.  ``t_step`` is ``self.dt_out`` / ``self.dt_in`` (or any sub-step of
        the propagator's time kind); ``eng`` must be this propagator's ``eng_out`` / ``eng_in`` view (the
        fused kernels evaluate those operators themselves) or None."""
        if eng is not None and not isinstance(eng, _OperatorView):
            return self._single_step_tables(t_step, eng)
        ts = complex(t_step)
        if self._time == 'imag':
            if abs(ts.real) > 0:
                raise ValueError("imaginary-time propagator: t_step must be -1j * dt")
            dt_sub = -ts.imag
        else:
            if abs(ts.imag) > 0:
                raise ValueError("real-time propagator: t_step must be real")
            dt_sub = ts.real
        self._plan.single_step(dt_sub)
        self._psik_cache = None

    def _single_step_tables(self, t_step, eng):
        """``single_step`` with caller-supplied operator tables ``eng = {'kin': [2], 'pot': [2], 'coupl': 2x2}`` of
        complex (Ny, Nx) tensors, as the reference applies them (tensor_propagator.py:242-271).  Not the fused path: the
        transforms and both normalisations run in the library's kernels (``sgpe_fft2d``, ``sgpe_normalise``), the
        table products are element-wise device operations."""
        if self._long:
            raise NotImplementedError("custom operator tables are not available on long-line meshes")
        dev = self._dev

        def table(t):
            t = t if isinstance(t, torch.Tensor) else torch.as_tensor(np.asarray(t))
            return t.to(device=dev, dtype=torch.complex128)

        kin = [table(k) for k in eng['kin']]
        pot = [table(p) for p in eng['pot']]
        psik = [kin[c] * self.psik[c] for c in range(2)]                                  # :242
        psi = ttools.ifft_2d(psik, self._dr)                                              # :243
        psi, dens = ttools.norm(psi, self._dv_r, self.atom_num)                           # :244
        g = self.g_sc
        int_eng = [g['uu'] * dens[0] + g['ud'] * dens[1], g['dd'] * dens[1] + g['ud'] * dens[0]]     # :247-248
        int_op = ttools.evolution_op(t_step / 2, int_eng)                                 # :249
        psi = [o * p for o, p in zip(int_op, psi)]                                        # :250
        if self.is_coupling:                                                              # :252-254
            cpl = [[table(e) for e in row] for row in eng['coupl']]
            psi = [cpl[r][0] * psi[0] + cpl[r][1] * psi[1] for r in range(2)]
        psi = [pot[c] * psi[c] for c in range(2)]                                         # :256
        if self.is_coupling:                                                              # :258-260
            psi = [cpl[r][0] * psi[0] + cpl[r][1] * psi[1] for r in range(2)]
        psi = [o * p for o, p in zip(int_op, psi)]                                        # :267
        psik = ttools.fft_2d(psi, self._dr)                                               # :269
        psik = [kin[c] * psik[c] for c in range(2)]                                       # :270
        psik, _ = ttools.norm(psik, self._dv_k, self.atom_num)                            # :271
        self.psik = psik

    def prop_loop(self, n_steps):
        """tensor_propagator.py:151-212 — step loop with per-step populations, optional sampling, final
        energy; returns a PropResult."""
        pop_times = np.linspace(0, self.n_steps * np.abs(self.t_step), n_steps)
        pops_dev = torch.zeros((1, max(n_steps, 1), 2), dtype=torch.float64, device=self._dev)
        track = {}
        if self.track_energy is not None:
            eng_dev = torch.zeros((1, max(n_steps, 1), 4), dtype=torch.float64, device=self._dev)
            track = dict(energy=eng_dev, kl_term=2 * self.kL_recoil * float(bool(self.is_coupling)),
                         unwrap=self.track_energy)
        done = 0
        if self.is_sampling:
            n_samples = int(n_steps / self.sample_rate)
            rate = int(self.sample_rate)
            shape = (n_samples, 2, self._plan.ny, self._plan.nx)
            sampled_host = torch.empty(shape, dtype=torch.complex128, pin_memory=True)
            sampled_times = np.linspace(0, self.n_steps * np.abs(self.t_step), n_samples)
            chunks = range(n_samples)
            if self._progress and _tqdm is not None:
                chunks = _tqdm(chunks)
            for idx in chunks:                               # sample BEFORE the step (:186-189)
                snap = self._plan.store()
                sampled_host[idx].copy_(snap[0].to(torch.complex128), non_blocking=True)
                self._plan.full_steps(rate, pops_dev, first=done, **track)
                done += rate
        if done < n_steps:
            self._plan.full_steps(n_steps - done, pops_dev, first=done, **track)
        self._psik_cache = None

        energy = self.eng_expect(None)
        vals = pops_dev[0, :n_steps].cpu().numpy().copy()
        pops = {'times': pop_times, 'vals': vals}

        if self.is_sampling:                                 # :198-204
            torch.cuda.synchronize(self._dev)
            test_name = self.paths['trial'] + 'psik_sampled'
            file_name = next_available_path(test_name, self.paths['folder'], '.npz')
            np.savez(file_name, psiks=sampled_host.numpy(), times=sampled_times)
        else:
            file_name = None

        psik_dev = self.psik
        if self._long:
            rs = self._plan.real_space()[0].to(torch.complex128)
            psi_dev = [rs[0], rs[1]]
        else:
            psi_dev = ttools.ifft_2d(psik_dev, self._dr)
        psik = _to_host(psik_dev)
        psi = _to_host(psi_dev)
        result = PropResult(psi, psik, energy, pops, file_name)
        if track:
            result.eng_history = track['energy'][0, :n_steps].cpu().numpy().copy()
        return result

    def eng_expect(self, psik=None, unwrap=None):
        """tensor_propagator.py:273-324 — [<total>, <kin>, <pot>, <int>] (raw grid sums), on the GPU.
        ``unwrap``: 'herraez' | 'none' | 'local' (default: the constructor's choice)."""
        unwrap = self.unwrap if unwrap is None else unwrap
        if psik is not None:
            assert len(psik) == 2, ("Requires two spinor components to calculate "
                                    "the energy expectation value.")
            comps = [torch.as_tensor(p) if not isinstance(p, torch.Tensor) else p for p in psik]
            psik = torch.stack([c.to(self._dev) for c in comps]).reshape(1, 2, self._plan.ny, self._plan.nx)
        kl_term = 2 * self.kL_recoil * float(bool(self.is_coupling))
        # the coupling energy uses the full coupling grid even where the step skips an all-zero one
        out = self._plan.energy(psik, kl_term=kl_term, unwrap=unwrap)
        return [float(v) for v in out[0].cpu().numpy()]

    def kin_expect_spectral(self, psik=None):
        """Kinetic energy per component from the k-space density, ``dv_k * sum kin_eng_spin[c] * abs(psik[c])**2``
        [hbar omega_x] — an alternative to the finite-difference / unwrapped-phase kinetic term of ``eng_expect``
        that needs no phase (no reference equivalent; the reference's number is a raw grid sum: multiply it by
        ``dv_r`` to compare)."""
        if psik is not None:
            comps = [torch.as_tensor(p) if not isinstance(p, torch.Tensor) else p for p in psik]
            psik = torch.stack([c.to(self._dev) for c in comps]).reshape(1, 2, self._plan.ny, self._plan.nx)
        if self._long:
            ksep = split_separable(np.stack(self._kin_np))
            if ksep is None:
                raise NotImplementedError("long-line meshes need a separable kinetic grid")
            return [float(v) for v in self._plan.kinetic_spectral(psik, ksep[0], ksep[1])[0].cpu().numpy()]
        return [float(v) for v in self._plan.kinetic_spectral(psik)[0].cpu().numpy()]
