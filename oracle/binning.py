"""Oracle (TEST INFRASTRUCTURE): event binning, numpy restatement.

Follows ``GEN1Dataset.slice_events`` (``yolox/data/datasets/gen1.py:313-328``) and
``GEN1Dataset.agrregate`` methods ``'sum'`` (``:333-349``) and ``'micro_sum'``
(``:355-360``) of the reference.  The same code is duplicated in ``gen4.py`` (720x1280)
and ``rvt_gen4.py:411-454``.

Semantics restated (integer, bit-exact target):
  * ``tw = (t[-1] - t[0]) // Tm``  (int64 floor division, overlap == 0)
  * micro-window k is ``[t0 + k*tw, t0 + (k+1)*tw)`` located with two *left* ``searchsorted``
    on the time-sorted stream; events with ``t >= t0 + Tm*tw`` are dropped; ``tw == 0``
    makes every window empty.
  * channel 0 counts ``p == 0``, channel 1 counts ``p != 0``; bin = ``y*W + x``.
  * output ``float64 [Tm, 2, H, W]`` (the reference's dtype); ``None``/empty -> zeros.
"""
from __future__ import annotations

import numpy as np


def slice_bounds(t: np.ndarray, num_slice: int):
    """(start, end) index arrays of the ``num_slice`` micro-windows (gen1.py:313-326)."""
    if len(t) <= 0:
        return None
    time_window = (t[-1] - t[0]) // num_slice
    window_start = np.arange(num_slice) * time_window + t[0]
    window_end = window_start + time_window
    return np.searchsorted(t, window_start), np.searchsorted(t, window_end)


def aggregate_sum(x, y, p, H: int, W: int) -> np.ndarray:
    """Per-polarity pixel histogram of one slice (gen1.py:333-349)."""
    frame = np.zeros((2, H * W))
    if x is None:
        return frame.reshape(2, H, W)
    xi = x.astype(int)
    yi = y.astype(int)
    off = p == 0
    for c, m in enumerate((off, np.logical_not(off))):
        counts = np.bincount(yi[m] * W + xi[m])
        frame[c][np.arange(counts.size)] += counts
    return frame.reshape(2, H, W)


def micro_sum(x, y, t, p, H: int, W: int, Tm: int) -> np.ndarray:
    """``agrregate(events, 'micro_sum')`` for one window (gen1.py:355-360)."""
    if t is None or len(t) == 0:
        return np.zeros((Tm, 2, H, W))
    start, end = slice_bounds(t, Tm)
    return np.stack([aggregate_sum(x[s:e], y[s:e], p[s:e], H, W) for s, e in zip(start, end)])


def micro_sum_batch(x, y, t, p, offsets, H: int, W: int, Tm: int) -> np.ndarray:
    """Batch of independent windows delimited by ``offsets[B+1]`` -> ``[B, Tm, 2, H, W]``.

    This is what the DataLoader collate (gen1.py:524-528) stacks with Tl == 1.
    """
    B = len(offsets) - 1
    out = np.zeros((B, Tm, 2, H, W))
    for b in range(B):
        s, e = int(offsets[b]), int(offsets[b + 1])
        out[b] = micro_sum(x[s:e], y[s:e], t[s:e], p[s:e], H, W, Tm)
    return out
