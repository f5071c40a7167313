"""Figures: matplotlib is not installed in this image, so these tests drive the plotting code against
a recording stand-in for the few matplotlib entry points it uses — they check what is plotted (arrays, extents, zoom
windows, colour limits), the file names and the return values, not matplotlib itself."""
import os
import sys
import tempfile
import types
from unittest import mock

import numpy as np
import pytest


class _Recorder:
    """Stands in for Figure / Axes / AxesImage / Colorbar: remembers every call, returns recorders."""

    def __init__(self, name, log):
        self._name, self._log = name, log

    def __getattr__(self, item):
        if item.startswith('__'):
            raise AttributeError(item)

        def call(*args, **kwargs):
            self._log.append((self._name + '.' + item, args, kwargs))
            if item == 'plot':
                return [_Recorder('line', self._log), _Recorder('line', self._log)]
            return _Recorder(item, self._log)
        return call


def fake_matplotlib(log):
    plt = types.ModuleType('matplotlib.pyplot')

    def subplots(rows, cols, **kw):
        log.append(('plt.subplots', (rows, cols), kw))
        axs = [_Recorder('ax', log) for _ in range(cols)]
        arr = np.empty(cols, dtype=object)
        arr[:] = axs
        return _Recorder('fig', log), (arr if cols > 1 else axs[0])

    def savefig(path):
        log.append(('plt.savefig', (path,), {}))
        open(path, 'w').close()

    plt.subplots = subplots
    plt.figure = lambda **kw: (log.append(('plt.figure', (), kw)), _Recorder('fig', log))[1]
    plt.savefig = savefig
    plt.show = lambda: log.append(('plt.show', (), {}))
    plt.tight_layout = lambda: None
    plt.close = lambda fig: log.append(('plt.close', (), {}))
    gridspec = types.ModuleType('matplotlib.gridspec')

    class GridSpec:
        def __init__(self, rows, cols):
            self.shape = (rows, cols)

        def __getitem__(self, key):
            return key
    gridspec.GridSpec = GridSpec
    animation = types.ModuleType('matplotlib.animation')

    class FuncAnimation:
        def __init__(self, fig, func, frames, blit=False):
            self.func, self.frames = func, frames

        def save(self, path, writer=None):
            for f in range(self.frames):
                self.func(f)
            open(path, 'w').close()
    animation.FuncAnimation = FuncAnimation
    animation.writers = {'ffmpeg': lambda fps, bitrate: ('ffmpeg', fps, bitrate)}
    mpl = types.ModuleType('matplotlib')
    mpl.pyplot, mpl.gridspec, mpl.animation = plt, gridspec, animation
    return {'matplotlib': mpl, 'matplotlib.pyplot': plt, 'matplotlib.gridspec': gridspec,
            'matplotlib.animation': animation}


@pytest.fixture
def ps():
    from spinor_gpe_b200 import PSpinor
    tmp = os.path.join(tempfile.mkdtemp(prefix='sgpe_plot_'), 'trial') + os.sep
    p = PSpinor(tmp, overwrite=True, mesh_points=(32, 64), r_sizes=(8, 16), atom_num=1e3)
    p.coupling_setup(wavel=790.1e-9, kin_shift=True)
    return p


def test_without_matplotlib_the_error_says_so(ps):
    with mock.patch.dict(sys.modules, {'matplotlib': None, 'matplotlib.pyplot': None}):
        with pytest.raises(ImportError, match='matplotlib'):
            ps.plot_rdens()


def test_pspinor_figures(ps):
    from spinor_gpe_b200 import tensor_tools as tt
    log = []
    with mock.patch.dict(sys.modules, fake_matplotlib(log)):
        ps.plot_rdens(scale=2.0)
        shown = [c for c in log if c[0] == 'ax.imshow']
        assert len(shown) == 2
        np.testing.assert_array_equal(shown[0][1][0], tt.density(ps.psi)[0])
        np.testing.assert_allclose(shown[0][2]['extent'], np.array([-8, 8, -16, 16]) / 2.0)
        log.clear()
        ps.plot_kdens(spin=1)
        shown = [c for c in log if c[0] == 'ax.imshow']
        assert len(shown) == 1
        np.testing.assert_array_equal(shown[0][1][0], tt.density(ps.psik)[1])
        k = ps.space['k_sizes']
        np.testing.assert_allclose(shown[0][2]['extent'], [-k[0], k[0], -k[1], k[1]])
        log.clear()
        ps.plot_rphase()
        assert len([c for c in log if c[0] == 'ax.imshow']) == 2
        with pytest.raises(AssertionError):
            ps.plot_rdens(spin=2)
        log.clear()
        fig, plots = ps.plot_spins(rscale=ps.rad_tf, kscale=ps.kL_recoil, zoom=4)
        assert set(plots) == {'r', 'ph', 'k'} and all(len(v) == 2 for v in plots.values())
        saved = [c[1][0] for c in log if c[0] == 'plt.savefig']
        assert saved == [ps.paths['data'] + 'spin_dens_phase1-trial.pdf']
        zoomed = [c for c in log if c[0] == 'add_subplot.set_xlim']
        np.testing.assert_allclose(zoomed[0][1][0], np.array([-k[0], k[0]]) / ps.kL_recoil / 4)
        ps.plot_spins()
        assert os.path.exists(ps.paths['data'] + 'spin_dens_phase2-trial.pdf')


def test_propresult_figures_and_movie(ps):
    from spinor_gpe_b200 import PropResult
    from spinor_gpe_b200 import tensor_tools as tt
    rng = np.random.default_rng(3)
    n = 6
    pops = {'times': np.linspace(0, 1, n), 'vals': 500 + rng.normal(size=(n, 2))}
    frames = np.array([[p * np.exp(0.1j * f) for p in ps.psik] for f in range(3)])
    sampled = ps.paths['trial'] + 'psik_sampled1-trial.npz'
    np.savez(sampled, psiks=frames, times=np.linspace(0, 1, 3))
    res = PropResult(ps.psi, ps.psik, [0.0] * 4, pops, sampled)
    res.paths, res.space, res.time_scale = ps.paths, ps.space, ps.time_scale
    log = []
    with mock.patch.dict(sys.modules, fake_matplotlib(log)):
        fig, plots = res.plot_total(kscale=ps.kL_recoil, zoom=2)
        assert set(plots) == {'r', 'ph', 'k'}
        total = [c for c in log if c[0] == 'add_subplot.imshow'][0][1][0]
        np.testing.assert_allclose(total, sum(tt.density(ps.psi)))
        assert os.path.exists(ps.paths['data'] + 'total_dens_phase1-trial.pdf')
        log.clear()
        res.plot_pops(scaled=True)
        line = [c for c in log if c[0] == 'add_subplot.plot'][0]
        np.testing.assert_allclose(line[1][0], pops['times'] * ps.time_scale)
        assert os.path.exists(ps.paths['data'] + 'pop_evolution1-trial.pdf')
        log.clear()
        res.make_movie(rscale=ps.rad_tf, kscale=ps.kL_recoil, norm_type='half')
        assert os.path.exists(ps.paths['data'] + 'prop_movie1-trial.mp4')
        updates = [c for c in log if c[0] == 'imshow.set_data']
        assert len(updates) == 3 * 6                                  # three frames, six images
        clim = [c for c in log if c[0] == 'imshow.set_clim'][0]
        dens0 = tt.density(tt.ifft_2d(list(frames[0]), ps.space['dr']))
        np.testing.assert_allclose(clim[1][1], sum(d.max() for d in dens0) / 2.0)
        res.sampled_path = None
        with pytest.warns(UserWarning):
            res.make_movie()


def test_progress_helpers(capsys):
    from spinor_gpe_b200 import plotting_tools as pt
    assert pt.time_remaining(10, 3610, 1.0) == '[01:00:00]'
    assert pt.time_remaining(0, 125, 1.0) == '[00:02:05]'
    pt.progress_message(0, 5)
    pt.progress_message(1, 5)
    assert '1/5' in capsys.readouterr().out
