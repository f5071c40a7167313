"""``PSpinor`` — problem definition for the quasi-2D pseudospin-1/2 GPE, API-compatible with the
reference's ``spinor_gpe/pspinor/pspinor.py`` (grids, trap, Thomas-Fermi state, Raman coupling, detuning,
``imaginary()`` / ``real()``).  Set-up is one-off host work in NumPy, as in the reference; the time
stepping it launches runs on the GPU (``tensor_propagator.TensorPropagator``).

All quantities are dimensionless in harmonic-oscillator units of the x trap frequency
(length a_x = sqrt(hbar / m omega_x), energy hbar omega_x, time 1/omega_x).
"""
import os
import shutil

import numpy as np
from scipy.ndimage import fourier_shift

from . import constants as const
from . import tensor_tools as ttools
from . import tensor_propagator as tprop
from . import plotting_tools as ptools

ROOT_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class PSpinor:
    """A pseudospin-1/2 condensate on an (Nx, Ny) mesh.  Arrays are (Ny, Nx), x contiguous."""

    # pylint: disable=too-many-instance-attributes,too-many-arguments
    def __init__(self, path, omeg=None, g_sc=None, mesh_points=(256, 256), r_sizes=(16, 16), atom_num=1e4,
                 pop_frac=(0.5, 0.5), **kwargs):
        phase_factor = kwargs.get('phase_factor', 1)
        overwrite = kwargs.get('overwrite', False)
        self.setup_data_path(path, overwrite)

        self.atom_num = atom_num
        self.space = {}
        assert sum(pop_frac) == 1.0, "Total population must equal 1."
        self.pop_frac = pop_frac

        if omeg is None:
            w0 = 2 * np.pi * 50
            omeg = {'x': w0, 'y': w0, 'z': 40 * w0}
        assert set(omeg.keys()) == {'x', 'y', 'z'}, "Keys for `omeg` must have the form: {'x', 'y', 'z'}."
        self.omeg = omeg
        if g_sc is None:
            g_sc = {'uu': 1.0, 'dd': 0.995, 'ud': 0.995}
        assert set(g_sc.keys()) == {'uu', 'dd', 'ud'}, "Keys for `g_sc` must have the form: {'uu', 'dd', 'ud'}."
        self.g_sc = dict(g_sc)

        self.compute_tf_params()
        self.compute_spatial_grids(mesh_points, r_sizes)
        self.compute_energy_grids()
        self.compute_tf_psi(phase_factor)
        self.no_coupling_setup()

        self.rand_seed = None
        self.prop = None
        shape = (int(mesh_points[1]), int(mesh_points[0]))
        self.coupling = np.zeros(shape)
        self.detuning = np.zeros(shape)
        self.rot_coupling = True

    # ------------------------------------------------------------------ directories
    def setup_data_path(self, path, overwrite):
        """Create <data>/, <data>/code/, <data>/trial_data/ (reference pspinor.py:198-246).  A relative
        ``path`` lands under <repo>/data/."""
        data = path if os.path.isabs(path) else os.path.join(ROOT_DIR, 'data', path)
        if os.path.isdir(data):
            if not overwrite:
                raise FileExistsError(f"The directory {data} already exists. To overwrite this directory, "
                                      "supply the parameter `overwrite=True`.")
            shutil.rmtree(data)
        data = os.path.normpath(data) + os.sep
        self.paths = {'data': data, 'trial': data + 'trial_data' + os.sep, 'code': data + 'code' + os.sep,
                      'folder': os.path.basename(os.path.normpath(data))}
        for key in ('data', 'code', 'trial'):
            os.makedirs(self.paths[key], exist_ok=True)

    # ------------------------------------------------------------------ scales and grids
    def compute_tf_params(self, species='Rb87'):
        """Thomas-Fermi scales (pspinor.py:283-313): a_x, a_sc, chemical potential, rescaled g_sc."""
        y_trap = self.omeg['y'] / self.omeg['x']
        z_trap = self.omeg['z'] / self.omeg['x']
        self.a_x = np.sqrt(const.hbar / (const.Rb87['m'] * self.omeg['x']))
        self.a_sc = const.Rb87['a_sc'] / self.a_x if species == 'Rb87' else 1
        self.chem_pot = (4 * self.atom_num * self.a_sc * y_trap * np.sqrt(z_trap / (2 * np.pi))) ** (1 / 2)
        g_scale = np.sqrt(8 * z_trap * np.pi) * self.a_sc
        self.g_sc.update({k: g_scale * v for k, v in self.g_sc.items()})
        self.rad_tf = np.sqrt(2 * self.chem_pot)
        self.time_scale = 1 / self.omeg['x']

    @classmethod
    def _compute_lin(cls, sizes, points, axis=0):
        return np.linspace(-sizes[axis], sizes[axis], num=points[axis], endpoint=False)

    def compute_spatial_grids(self, mesh_points=(256, 256), r_sizes=(16, 16)):
        """Real- and momentum-space meshes (pspinor.py:315-364)."""
        assert all(p % 2 == 0 for p in mesh_points), f"Number of mesh points {mesh_points} should be powers of 2."
        mesh_points = np.array(mesh_points)
        r_sizes = np.array(r_sizes)
        sp = self.space
        sp['dr'] = 2 * r_sizes / mesh_points
        k_sizes = np.pi / sp['dr']
        sp['dk'] = np.pi / r_sizes
        sp['x'] = self._compute_lin(r_sizes, mesh_points, axis=0)
        sp['y'] = self._compute_lin(r_sizes, mesh_points, axis=1)
        sp['kx'] = self._compute_lin(k_sizes, mesh_points, axis=0)
        sp['ky'] = self._compute_lin(k_sizes, mesh_points, axis=1)
        sp['x_mesh'], sp['y_mesh'] = np.meshgrid(sp['x'], sp['y'])
        sp['kx_mesh'], sp['ky_mesh'] = np.meshgrid(sp['kx'], sp['ky'])
        sp['dv_r'] = np.prod(sp['dr'])
        sp['dv_k'] = np.prod(sp['dk'])
        sp['mesh_points'] = mesh_points
        sp['r_sizes'] = r_sizes
        sp['k_sizes'] = k_sizes

    @property
    def pot_eng(self):
        """2D potential energy grid [hbar omega_x]."""
        return self._pot_eng

    @pot_eng.setter
    def pot_eng(self, array):
        self._pot_eng = array
        self.pot_eng_spin = [self._pot_eng] * 2

    @property
    def kin_eng(self):
        """2D kinetic energy grid [hbar omega_x]."""
        return self._kin_eng

    @kin_eng.setter
    def kin_eng(self, array):
        self._kin_eng = array
        self.kin_eng_spin = [self._kin_eng] * 2

    def compute_energy_grids(self):
        """Harmonic trap and free-particle dispersion (pspinor.py:413-430)."""
        y_trap = self.omeg['y'] / self.omeg['x']
        self.pot_eng = (self.space['x_mesh'] ** 2 + (y_trap * self.space['y_mesh']) ** 2) / 2
        self.kin_eng = (self.space['kx_mesh'] ** 2 + self.space['ky_mesh'] ** 2) / 2

    def compute_tf_psi(self, phase_factor=1.0):
        """Thomas-Fermi initial state, its FFT and healing lengths (pspinor.py:248-281)."""
        assert abs(phase_factor) == 1.0, "Relative phase factor must have unit magnitude."
        g_bare = [self.g_sc['uu'], self.g_sc['dd']]
        profile = np.real(np.sqrt((self.chem_pot - self.pot_eng + 0.j)))
        self.psi = [profile * np.sqrt(pop / abs(g)) for pop, g in zip(self.pop_frac, g_bare)]
        self.psi[1] = self.psi[1] * phase_factor
        self.psi, _ = ttools.norm(self.psi, self.space['dv_r'], self.atom_num)
        self.psik = ttools.fft_2d(self.psi, self.space['dr'])
        with np.errstate(divide='ignore'):
            self.heal = [(8 * np.pi * np.max(np.abs(p) ** 2) * self.a_sc) ** (-1 / 2) for p in self.psi]
        np.savez(self.paths['trial'] + 'tf_wf-' + self.paths['folder'], psi=self.psi, psik=self.psik)

    def _calc_atoms(self, psi=None, space='r'):
        """Total atom number of ``psi`` (default: the current state) in 'r' or 'k' space."""
        if space == 'r':
            psi, vol = (self.psi if psi is None else psi), self.space['dv_r']
        elif space == 'k':
            psi, vol = (self.psik if psi is None else psi), self.space['dv_k']
        else:
            raise ValueError("space must be 'r' or 'k'")
        return ttools.calc_atoms(psi, vol)

    # ------------------------------------------------------------------ coupling
    def no_coupling_setup(self):
        """Defaults without Raman coupling (pspinor.py:462-467)."""
        self.is_coupling = False
        self.kL_recoil = 1.0          # pylint: disable=invalid-name
        self.EL_recoil = 1.0          # pylint: disable=invalid-name

    def coupling_setup(self, wavel=790.1e-9, scale=1.0, kin_shift=False):
        """Raman recoil units and (optionally) the spin-dependent kinetic shift (pspinor.py:469-501)."""
        self.is_coupling = True
        self.kL_recoil = scale * (np.sqrt(2) * np.pi / wavel * self.a_x)
        self.EL_recoil = self.kL_recoil ** 2 / 2
        shift = self.space['kx_mesh'] * self.kL_recoil if kin_shift else 0
        self.kin_eng_spin = [self.kin_eng + shift, self.kin_eng - shift]
        self.kin_eng_spin = [k - np.min(k) for k in self.kin_eng_spin]

    def shift_momentum(self, psik=None, scale=1.0, frac=(0.5, 0.5)):
        """Move fractions of each component's momentum peak by -/+ scale*kL (pspinor.py:503-545)."""
        assert self.is_coupling, (f"The `is_coupling` option is {self.is_coupling}. "
                                  "Initialize coupling with `coupling_setup()`.")
        if psik is None:
            psik = self.psik
        shift = scale * self.kL_recoil / self.space['dk'][0]
        spectrum = ttools.fft_2d(psik, self.space['dr'])
        moved = []
        for comp in spectrum:
            plus = fourier_shift(comp, shift=[0, shift], axis=1)
            minus = fourier_shift(comp, shift=[0, -shift], axis=1)
            moved.append(frac[0] * plus + frac[1] * minus)
            frac = np.flip(frac)
        self.psik = ttools.ifft_2d(moved, self.space['dr'])
        self.psi = ttools.ifft_2d(self.psik, self.space['dr'])

    @property
    def coupling(self):
        """2D Raman coupling grid [hbar omega_x]."""
        return self._coupling

    @coupling.setter
    def coupling(self, array):
        self._coupling = array

    @property
    def detuning(self):
        """2D detuning grid [hbar omega_x]; setting it rebuilds pot_eng_spin (pspinor.py:570-575)."""
        return self._detuning

    @detuning.setter
    def detuning(self, array):
        self._detuning = array
        self.pot_eng_spin = [self.pot_eng + self._detuning / 2, self.pot_eng - self._detuning / 2]

    def _axis_mesh(self, axis):
        if axis == 0:
            return self.space['x_mesh']
        if axis == 1:
            return self.space['y_mesh']
        raise ValueError("axis must be 0 (x) or 1 (y)")

    def coupling_grad(self, slope, offset, axis=1):
        """Linear coupling gradient (pspinor.py:577-606)."""
        self.coupling = self._axis_mesh(axis) * slope + offset

    def coupling_uniform(self, value):
        """Uniform coupling (pspinor.py:608-625)."""
        assert value >= 0, f"Cannot have a negative coupling value: {value}."
        self.coupling = np.ones_like(self.space['x_mesh']) * value

    def detuning_grad(self, slope, offset=0.0, axis=1):
        """Linear detuning gradient (pspinor.py:627-660)."""
        self.detuning = self._axis_mesh(axis) * slope + offset

    def detuning_uniform(self, value):
        """Uniform detuning (pspinor.py:662-678)."""
        self.detuning = np.ones_like(self.space['x_mesh']) * value

    def seed_vortices(self, positions, windings):
        """Imprint vortices (core profile r/sqrt(r^2+1) in healing lengths, phase winding) at
        ``positions`` (pspinor.py:680-745)."""
        positions = np.array(positions)
        assert positions.shape[-1] == 2, "Positions should be ordered pairs, e.g. (x, y)."
        same = False
        if positions.ndim == 2:
            n_pos = len(positions)
            positions = np.array([positions, positions])
            same = True
        else:
            assert positions.ndim == 3, "`positions` must be at most a three-dimensional array."
        windings = np.array(windings)
        if windings.shape == (1,):
            windings = windings * np.ones(positions.shape[:-1], dtype=windings.dtype)
        elif windings.ndim == 1:
            assert same and len(windings) == n_pos, \
                "The number of supplied windings must match the number of supplied positions"
            windings = np.array([windings, windings])
        else:
            assert windings.shape == positions.shape[:-1], \
                "The number of supplied windings must match the number of supplied positions"
        for i in range(2):
            for j, (x0, y0) in enumerate(positions[i]):
                dx = self.space['x_mesh'] - x0
                dy = self.space['y_mesh'] - y0
                rho = np.sqrt(dx ** 2 + dy ** 2) / self.heal[i]
                self.psi[i] = self.psi[i] * (rho / np.sqrt(rho ** 2 + 1)) \
                    * np.exp(windings[i, j] * 1j * np.arctan2(dy, dx))
        self.psik = ttools.fft_2d(self.psi, delta_r=self.space['dr'])

    def seed_regular_vortices(self):
        raise NotImplementedError()          # pspinor.py:747-752

    def seed_random_vortices(self, N):       # pylint: disable=invalid-name
        raise NotImplementedError()          # pspinor.py:754-756

    # ------------------------------------------------------------------ figures (host-side, need matplotlib)
    def plot_rdens(self, psi=None, spin=None, cmap='viridis', scale=1.0):
        """Real-space density (pspinor.py:758-787)."""
        ptools.plot_dens(self.psi if psi is None else psi, spin, cmap, scale,
                         extent=ptools.extents_of(self.space, rscale=scale)['r'])

    def plot_kdens(self, psik=None, spin=None, cmap='viridis', scale=1.0):
        """Momentum-space density (pspinor.py:789-818)."""
        ptools.plot_dens(self.psik if psik is None else psik, spin, cmap, scale,
                         extent=ptools.extents_of(self.space, kscale=scale)['k'])

    def plot_rphase(self, psi=None, spin=None, cmap='twilight_shifted', scale=1.0):
        """Real-space phase (pspinor.py:820-851)."""
        ptools.plot_phase(self.psi if psi is None else psi, spin, cmap, scale,
                          extent=ptools.extents_of(self.space, rscale=scale)['r'])

    def plot_spins(self, rscale=1.0, kscale=1.0, cmap='viridis', save=True, ext='.pdf', zoom=1.0):
        """Densities and phases of both components (pspinor.py:853-888); returns (fig, all_plots)."""
        return ptools.plot_spins(self.psi, self.psik, ptools.extents_of(self.space, rscale, kscale), self.paths,
                                 cmap=cmap, save=save, ext=ext, zoom=zoom)

    # ------------------------------------------------------------------ propagation
    def _propagate(self, time, t_step, n_steps, device, is_sampling, n_samples, **kw):
        prop = tprop.TensorPropagator(self, t_step, n_steps, device, time=time, is_sampling=is_sampling,
                                      n_samples=n_samples, **kw)
        result = prop.prop_loop(prop.n_steps)
        result.paths = self.paths
        result.time_scale = self.time_scale
        result.space = self.space
        self.psik = result.psik              # chaining: the next run continues from here (pspinor.py:923-924)
        self.psi = result.psi
        return result, prop

    def imaginary(self, t_step, n_steps=1000, device='cuda', is_sampling=False, n_samples=1, **kw):
        """Imaginary-time propagation; returns (PropResult, TensorPropagator) (pspinor.py:890-925)."""
        return self._propagate('imag', t_step, n_steps, device, is_sampling, n_samples, **kw)

    def real(self, t_step, n_steps=1000, device='cuda', is_sampling=False, n_samples=1, **kw):
        """Real-time propagation; returns (PropResult, TensorPropagator) (pspinor.py:927-962)."""
        return self._propagate('real', t_step, n_steps, device, is_sampling, n_samples, **kw)
