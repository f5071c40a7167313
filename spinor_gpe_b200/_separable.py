"""Host-side detection of separable operator grids (one-off set-up, NumPy).

A grid g[c, y, x] is separable when g[c, y, x] = gx[c, x] + gy[c, y]; that holds for everything the
reference's PSpinor builds by default: the harmonic trap (pspinor.py:426-428), the free and Raman-shifted
dispersions (:429-430, :496-501) and uniform / linear-gradient detunings (:574-575, :660, :678).  For such
grids the kernels multiply by products of 1-D factor tables instead of evaluating exp / sincos per point.
"""
import os

import numpy as np

RTOL = 1e-13


def split_separable(grids, rtol=RTOL):
    """(2, Ny, Nx) -> (gx (2, Nx), gy (2, Ny)) with grids == gx[:, None, :] + gy[:, :, None] to within
    ``rtol`` of the largest entry, or None when the grids are not separable.

    The split is anchored at the row where the y-dependence is smallest, so that gy >= 0 with min(gy) = 0 and
    gx carries the rest: the factor tables exp(-gx tau), exp(-gy tau) then stay as well scaled as the
    operator itself (anchoring at a corner would make one table overflow in imaginary time on fine meshes,
    where k_max^2 tau / 2 exceeds 709)."""
    # (a list of per-component grids is taken as it is: stacking two 2048^2 grids costs more than the whole check)
    g = [np.asarray(c, dtype=np.float64) for c in grids]
    nc, (ny, nx) = len(g), g[0].shape
    gx = np.empty((nc, nx))
    gy = np.empty((nc, ny))
    for c in range(nc):
        col = g[c][:, 0]
        i0 = int(np.argmin(col))
        gy[c] = col - col[i0]
        gx[c] = g[c][i0, :]
    # a cheap look at a coarse sub-grid rejects most non-separable grids before the full check
    sy, sx = max(1, ny // 64), max(1, nx // 64)
    for c in range(nc):
        coarse = g[c][::sy, ::sx]
        scale = max(float(np.abs(coarse).max()), 1e-300)
        if np.abs(coarse - (gx[c, None, ::sx] + gy[c, ::sy, None])).max() > rtol * scale * 4:
            return None
    # full check in cache-sized row blocks (no full-size temporaries), a few host threads (NumPy releases the GIL)
    rows = max(1, (1 << 19) // max(1, nx))
    blocks = [(c, r) for c in range(nc) for r in range(0, ny, rows)]

    def check(job):
        c, r = job
        blk = g[c][r:r + rows]
        return (float(np.abs(blk - gx[c] - gy[c, r:r + rows, None]).max()), float(np.abs(blk).max()))

    if len(blocks) > 4:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
            res = list(pool.map(check, blocks))
    else:
        res = [check(b) for b in blocks]
    err = max(r[0] for r in res)
    if err <= rtol * max(max(r[1] for r in res), 1e-300):
        return gx, gy
    return None
