"""Oracle (TEST INFRASTRUCTURE): alternative event representations, numpy restatements.

``to_voxel_grid`` follows ``to_voxel_grid_numpy`` (``yolox/utils/event_reps.py:30-89``, adapted from Tonic): event
volume with bilinear interpolation in time.  PINNED by ``tests/golden/voxel.npz`` (the reference function run on
structured arrays of its own ``events_struct`` dtype and of a signed-polarity dtype).
"""
from __future__ import annotations

import numpy as np


def to_voxel_grid(x, y, t, p, H: int, W: int, n_bins: int = 10, p_is_bool: bool = True) -> np.ndarray:
    """One window -> ``float64 [n_bins, 1, H, W]``.  ``p_is_bool``: the reference's ``events_struct`` has a bool
    polarity field, so its ``pols[pols == 0] = -1`` (event_reps.py:62-63) stores True and every event weighs +1;
    with a signed integer field the weights are +1 / -1 as the comment there says."""
    n = len(x)
    if n == 0:
        return np.zeros((n_bins, 1, H, W), float)                                   # :46-47
    grid = np.zeros((n_bins, H, W), float).ravel()
    tt = np.asarray(t)
    ts = n_bins * (tt.astype(float) - tt[0]) / (tt[-1] - tt[0])                      # :52-56
    xs, ys = np.asarray(x).astype(int), np.asarray(y).astype(int)
    pols = np.ones(n) if p_is_bool else np.where(np.asarray(p) == 0, -1.0, 1.0)     # :61-62
    tis = ts.astype(int)
    dts = ts - tis
    left, right = pols * (1.0 - dts), pols * dts                                    # :64-67
    v = tis < n_bins
    np.add.at(grid, xs[v] + ys[v] * W + tis[v] * W * H, left[v])                    # :69-76
    v = (tis + 1) < n_bins
    np.add.at(grid, xs[v] + ys[v] * W + (tis[v] + 1) * W * H, right[v])             # :78-85
    return grid.reshape(n_bins, 1, H, W)
