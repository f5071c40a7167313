"""``PropResult`` — container returned by ``PSpinor.imaginary()/real()`` (reference
spinor_gpe/pspinor/prop_result.py:50-88, 304-328).  Plotting needs matplotlib, which is optional here."""
import numpy as np

from . import tensor_tools as ttools


def _need_matplotlib():
    try:
        import matplotlib  # noqa: F401
    except ImportError as exc:      # pragma: no cover
        raise ImportError("plotting needs matplotlib, which is not installed in this environment; "
                          "the numerical results (psi, psik, pops, eng_final, dens, phase) do not") from exc


class PropResult:
    """Results of a propagation: final wavefunctions, energies, per-step populations, sample file."""

    def __init__(self, psi_final, psik_final, eng_final, pops, sampled_path=None):
        self.psi = psi_final
        self.psik = psik_final
        self.eng_final = eng_final
        self.pops = pops
        self.sampled_path = sampled_path

        self.dens = ttools.density(self.psi)
        self.densk = ttools.density(self.psik)
        self.phase = ttools.phase(self.psi, uwrap=False, dens=self.dens)

        self.paths = dict()
        self.time_scale = None
        self.space = dict()

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

    def plot_spins(self, *args, **kwargs):
        _need_matplotlib()
        raise NotImplementedError("figure generation is outside the propagator path (SURVEY.md 8f-4)")

    plot_total = plot_pops = make_movie = plot_spins

    def plot_eng(self):
        raise NotImplementedError()          # a stub in the reference as well (prop_result.py:165-167)

    def analyze_vortex(self):
        raise NotImplementedError()          # prop_result.py:216-218
