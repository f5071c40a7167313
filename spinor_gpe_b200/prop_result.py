"""``PropResult`` — container returned by ``PSpinor.imaginary()/real()`` (reference
spinor_gpe/pspinor/prop_result.py:50-88, 304-328).  The figures need matplotlib, which is imported only when one is
requested (``plotting_tools``)."""
import os
import sys

import numpy as np

from . import plotting_tools as ptools
from . import tensor_tools as ttools


class PropResult:
    """Results of a propagation: final wavefunctions, energies, per-step populations, sample file."""

    def __init__(self, psi_final, psik_final, eng_final, pops, sampled_path=None):
        self.psi = psi_final
        self.psik = psik_final
        self.eng_final = eng_final
        self.pops = pops
        self.sampled_path = sampled_path
        self.eng_history = None      # (n_steps, 4) when the propagator tracked the energy of every step (extension)

        # densities and phase (prop_result.py:68-72) are evaluated when first read: at 2048^2 they cost more host time
        # than the whole propagation takes on the GPU
        self._dens = self._densk = self._phase = None

        self.paths = dict()
        self.time_scale = None
        self.space = dict()

    @property
    def dens(self):
        """Real-space densities of both components (prop_result.py:68)."""
        if self._dens is None:
            self._dens = ttools.density(self.psi)
        return self._dens

    @dens.setter
    def dens(self, value):
        self._dens = value

    @property
    def densk(self):
        """Momentum-space densities (prop_result.py:69)."""
        if self._densk is None:
            self._densk = ttools.density(self.psik)
        return self._densk

    @densk.setter
    def densk(self, value):
        self._densk = value

    @property
    def phase(self):
        """Wrapped phases, zeroed where the density is below 1e-6 of its maximum (prop_result.py:70-72)."""
        if self._phase is None:
            self._phase = ttools.phase(self.psi, uwrap=False, dens=self.dens)
        return self._phase

    @phase.setter
    def phase(self, value):
        self._phase = value

    def calc_separation(self):
        """Phase separation 1 - <n0 n1> / sqrt(<n0^2><n1^2>) (prop_result.py:84-88)."""
        overlap = np.sum(self.dens[0] * self.dens[1])
        return 1 - overlap / np.sqrt(np.sum(self.dens[0] ** 2) * np.sum(self.dens[1] ** 2))

    def rebin(self, arr, new_shape=(256, 256)):
        """Average-pool ``arr`` down to ``new_shape`` (prop_result.py:304-328)."""
        assert arr[0].shape == arr[1].shape
        if not tuple(new_shape) < tuple(arr[0].shape):
            return arr
        ny, nx = new_shape
        return [a.reshape(ny, a.shape[0] // ny, nx, a.shape[1] // nx).mean(-1).mean(1) for a in arr]

    def plot_spins(self, rscale=1.0, kscale=1.0, cmap='viridis', save=True, ext='.pdf', show=True, zoom=1.0):
        """Densities and phases of both components (prop_result.py:90-125); returns (fig, all_plots)."""
        return ptools.plot_spins(self.psi, self.psik, ptools.extents_of(self.space, rscale, kscale), self.paths,
                                 cmap=cmap, save=save, ext=ext, show=show, zoom=zoom)

    def plot_total(self, rscale=1.0, kscale=1.0, cmap='viridis', save=True, ext='.pdf', show=True, zoom=1.0):
        """Total densities and phase (prop_result.py:127-163); returns (fig, all_plots)."""
        return ptools.plot_total(self.psi, self.psik, ptools.extents_of(self.space, rscale, kscale), self.paths,
                                 cmap=cmap, save=save, ext=ext, show=show, zoom=zoom)

    def plot_pops(self, scaled=True, save=True, ext='.pdf'):
        """Populations of both components against time, and the size of their step-to-step change on a log scale
        (prop_result.py:169-214).  Saved as ``<data>/pop_evolution<i>-<folder><ext>``."""
        plt = ptools._pyplot()
        unit, label = (self.time_scale, 'Time [s]') if scaled else (1.0, 'Time [$1/\\omega_x$]')
        times = self.pops['times'] * unit
        fig = plt.figure(figsize=(12, 4))
        left, right = fig.add_subplot(121), fig.add_subplot(122)
        lines = left.plot(times, self.pops['vals'])
        left.legend(lines, ('Pop. $| \\uparrow\\rangle$', 'Pop. $| \\downarrow\\rangle$'))
        left.set_ylabel('Population')
        right.plot(times, np.abs(np.diff(self.pops['vals'])))
        right.set_ylabel('Abs. Population Difference')
        right.set_yscale('log')
        right.set_ylim(2e-16, None)
        for ax in (left, right):
            ax.set_xlabel(label)
            ax.grid(alpha=0.5)
        if save:
            plt.savefig(ptools.next_available_path(self.paths['data'] + 'pop_evolution', self.paths['folder'], ext))
        plt.show()

    def make_movie(self, rscale=1.0, kscale=1.0, cmap='viridis', play=False, zoom=1.0, norm_type='all'):
        """Animate the sampled wavefunctions (``psik_sampled*.npz`` written by ``prop_loop``) in the layout of
        ``plot_spins``; saved as ``<data>/prop_movie<i>-<folder>.mp4`` through matplotlib's ffmpeg writer
        (prop_result.py:220-302).  ``norm_type``: colour scale from the summed maxima ('all') or half of that."""
        import subprocess
        import warnings
        if not os.path.exists(str(self.sampled_path)):
            warnings.warn("Cannot generate propagation movie. No sampled wavefuntion data exists.")
            return
        divisor = {'all': 1.0, 'half': 2.0}[norm_type]
        from matplotlib import animation
        plt = ptools._pyplot()
        with np.load(self.sampled_path) as sampled:
            psiks = sampled['psiks']
        fig, images = self.plot_spins(rscale, kscale, cmap, save=False, show=False, zoom=zoom)

        def draw(frame):
            psik = list(psiks[frame])
            psi = ttools.ifft_2d(psik, self.space['dr'])
            dens, densk = ttools.density(psi), ttools.density(psik)
            for key, data in (('r', dens), ('ph', ttools.phase(psi, uwrap=False, dens=dens)), ('k', densk)):
                for img, arr in zip(images[key], data):
                    img.set_data(arr)
            for key, data in (('r', dens), ('k', densk)):
                top = sum(np.max(d) for d in data) / divisor
                for img in images[key]:
                    img.set_clim(0, top)
            ptools.progress_message(frame, len(psiks))

        movie = animation.FuncAnimation(fig, draw, frames=len(psiks), blit=False)
        file_name = ptools.next_available_path(self.paths['data'] + 'prop_movie', self.paths['folder'], '.mp4')
        movie.save(file_name, writer=animation.writers['ffmpeg'](fps=5, bitrate=-1))
        plt.close(fig)
        if play:
            if sys.platform == 'win32':
                os.startfile(file_name)          # pylint: disable=no-member
            else:
                subprocess.call(['open' if sys.platform == 'darwin' else 'xdg-open', file_name])

    def plot_eng(self):
        raise NotImplementedError()          # a stub in the reference as well (prop_result.py:165-167)

    def analyze_vortex(self):
        raise NotImplementedError()          # prop_result.py:216-218
