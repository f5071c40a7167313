"""Figures of spinor wavefunctions — the function set of the reference's ``spinor_gpe/pspinor/plotting_tools.py``
(``plot_dens`` :104, ``plot_phase`` :150, ``plot_spins`` :197, ``plot_total`` :297, ``next_available_path`` :13,
``progress_message`` :40, ``time_remaining`` :70) with the same arguments, return values and output file names, so
that scripts written against the reference (its examples call these after every run) keep working.

Pure host-side visualisation of NumPy arrays, outside the propagator path.  matplotlib is imported
when a figure is requested, not at package import: the numerical API works without it, and asking for a figure
without it raises an ImportError that says so.
"""
import os
import sys
import time

import numpy as np

from . import tensor_tools as ttools

_PI_TICKS = (np.linspace(-np.pi, np.pi, 5), ['$-\\pi$', '', '$0$', '', '$\\pi$'])
_last_tick = None


def _pyplot():
    try:
        from matplotlib import pyplot as plt
    except ImportError as exc:
        raise ImportError("plotting needs matplotlib, which is not installed in this environment; the numerical "
                          "results (psi, psik, pops, eng_final, dens, phase) do not") from exc
    return plt


def next_available_path(file_name, trial_name, ext=''):
    """First ``<file_name><i>-<trial_name><ext>``, i = 1, 2, ..., that does not exist yet."""
    idx = 1
    while os.path.exists(f'{file_name}{idx}-{trial_name}{ext}'):
        idx += 1
    return f'{file_name}{idx}-{trial_name}{ext}'


def time_remaining(frame, n_total, its):
    """``[hh:mm:ss]`` left for ``n_total - frame`` iterations at ``its`` iterations per second."""
    left = (n_total - frame) / its
    hours = int(left // 3600)
    minutes = int((left - hours * 3600) // 60)
    seconds = int(np.mod(left, 60))
    return f'[{hours:02d}:{minutes:02d}:{seconds:02d}]'


def progress_message(frame, n_total):
    """tqdm-like one-line progress for loops tqdm cannot wrap (the animation writer); returns the time stamp."""
    global _last_tick
    now = time.time()
    if frame != 0 and _last_tick is not None:
        rate = 1 / max(now - _last_tick, 1e-9)
        sys.stdout.write(f'\r{frame}/{n_total}, {time_remaining(frame, n_total, rate)}, {rate:.2f} it/sec')
        sys.stdout.flush()
    _last_tick = now
    return _last_tick


def _pick(psi, spin):
    if spin is None:
        return list(psi)
    assert spin in (0, 1), f"The `spin` parameter should be 0 or 1, not {spin}."
    return [psi[spin]]


def _row_of_axes(n):
    fig, axs = _pyplot().subplots(1, n, sharex=True, sharey=True)
    return fig, list(np.atleast_1d(axs))


def plot_dens(psi, spin=None, cmap='viridis', scale=1., extent=None):
    """Density of one (``spin`` = 0 / 1) or both (None) components, real or momentum space."""
    comps = _pick(psi, spin)
    _, axs = _row_of_axes(len(comps))
    for ax, dens in zip(axs, ttools.density(comps)):
        ax.imshow(dens, cmap=cmap, extent=extent)
    _pyplot().show()


def plot_phase(psi, spin=None, cmap='twilight_shifted', scale=1, extent=None):
    """Phase of one or both components, zeroed where the density is below 1e-6 of its maximum."""
    comps = _pick(psi, spin)
    dens = ttools.density(comps)
    _, axs = _row_of_axes(len(comps))
    for ax, phz in zip(axs, ttools.phase(comps, uwrap=False, dens=dens)):
        ax.imshow(phz, cmap=cmap, extent=extent)
    _pyplot().show()


def _panel(fig, ax, data, extent, labels, cmap, phase=False, equal=False):
    """One image with its colour bar; densities start at 0, phases span [-pi, pi] with pi ticks."""
    kw = dict(cmap='twilight_shifted', vmin=-np.pi, vmax=np.pi) if phase else dict(cmap=cmap, vmin=0)
    if equal:
        kw['aspect'] = 'equal'
    img = ax.imshow(data, origin='lower', extent=extent, **kw)
    bar = fig.colorbar(img, ax=ax)
    if phase:
        bar.set_ticks(_PI_TICKS[0])
        bar.set_ticklabels(_PI_TICKS[1])
    ax.set_xlabel(labels[0])
    ax.set_ylabel(labels[1])
    return img


def _finish(fig, stem, paths, ext, save, show):
    plt = _pyplot()
    plt.tight_layout()
    if save:
        plt.savefig(next_available_path(paths['data'] + stem, paths['folder'], ext))
    if show:
        plt.show()


def plot_spins(psi, psik, extents, paths, cmap='viridis', save=True, ext='.pdf', show=True, zoom=1.0):
    """Six panels: real-space density, phase and momentum-space density of both components.  Saved as
    ``<data>/spin_dens_phase<i>-<folder><ext>``.  Returns ``(fig, {'r': [...], 'ph': [...], 'k': [...]})`` with the
    two ``AxesImage`` of each row (what ``PropResult.make_movie`` updates frame by frame)."""
    from matplotlib import gridspec
    dens = ttools.density(psi)
    rows = {'r': (dens, extents['r'], ('$x$', '$y$'), False),
            'ph': (ttools.phase(psi, uwrap=False, dens=dens), extents['r'], ('$x$', '$y$'), True),
            'k': (ttools.density(psik), extents['k'], ('$k_x$', '$k_y$'), False)}
    fig = _pyplot().figure(figsize=(5.5, 6.4))
    grid = gridspec.GridSpec(6, 4)
    images = {}
    for r, (key, (pair, extent, labels, is_phase)) in enumerate(rows.items()):
        axs = [fig.add_subplot(grid[2 * r:2 * r + 2, 2 * c:2 * c + 2]) for c in range(2)]
        images[key] = [_panel(fig, ax, d, extent, labels, cmap, phase=is_phase, equal=True) for ax, d in zip(axs, pair)]
        if key == 'k':
            window = np.asarray(extents['k']) / zoom
            for ax in axs:
                ax.set_xlim(window[:2])
                ax.set_ylim(window[2:])
    _finish(fig, 'spin_dens_phase', paths, ext, save, show)
    return fig, images


def plot_total(psi, psik, extents, paths, cmap='viridis', save=True, ext='.pdf', show=True, zoom=1.0):
    """Three panels: total real-space density, phase of psi_up + psi_down, total momentum-space density.  Saved as
    ``<data>/total_dens_phase<i>-<folder><ext>``.  Returns ``(fig, {'r': img, 'ph': img, 'k': img})``."""
    from matplotlib import gridspec
    dens_r = sum(ttools.density(psi))
    fig = _pyplot().figure()
    grid = gridspec.GridSpec(4, 4)
    ax_r, ax_ph, ax_k = fig.add_subplot(grid[0:2, 0:2]), fig.add_subplot(grid[0:2, 2:]), fig.add_subplot(grid[2:, 1:3])
    images = {'r': _panel(fig, ax_r, dens_r, extents['r'], ('$x$', '$y$'), cmap),
              'ph': _panel(fig, ax_ph, ttools.phase(sum(psi), uwrap=False, dens=dens_r), extents['r'], ('$x$', '$y$'),
                           cmap, phase=True),
              'k': _panel(fig, ax_k, sum(ttools.density(psik)), extents['k'], ('$k_x$', '$k_y$'), cmap)}
    window = np.asarray(extents['k']) / zoom
    ax_k.set_xlim(window[:2])
    ax_k.set_ylim(window[2:])
    _finish(fig, 'total_dens_phase', paths, ext, save, show)
    return fig, images


def extents_of(space, rscale=1.0, kscale=1.0):
    """``{'r': [x_min, x_max, y_min, y_max] / rscale, 'k': ... / kscale}`` from a PSpinor ``space`` dictionary."""
    def box(sizes, scale):
        sizes = np.asarray(sizes, dtype=float)
        return np.ravel(np.vstack((-sizes, sizes)).T) / scale
    return {'r': box(space['r_sizes'], rscale), 'k': box(space['k_sizes'], kscale)}
