"""Synthetic event streams: the only data source on the GPU box.

One "window" is a 50 ms slice of a time-sorted event stream ``(x:i16, y:i16, t:i64 us, p:u8)``,
the SoA form of the reference's ``events_struct`` (``yolox/utils/util.py:119-121``).  A batch is
the concatenation of B windows plus ``offsets[B+1]``.
"""
from __future__ import annotations

import numpy as np

GEN1 = (240, 304)       # H, W  (yolox/data/datasets/gen1.py img_size default)
MPX = (720, 1280)       # gen4.py


def make_window(rng: np.random.Generator, n: int, H: int, W: int, span_us: int = 50_000,
                hot_frac: float = 0.10, hot_sigma: float = 8.0):
    """n events: uniform background + a Gaussian "hot cluster" (atomic contention), sorted in t."""
    n = int(n)
    t = np.sort(rng.integers(0, span_us, n, dtype=np.int64))
    x = rng.integers(0, W, n, dtype=np.int64)
    y = rng.integers(0, H, n, dtype=np.int64)
    n_hot = int(n * hot_frac)
    if n_hot > 0:
        cx, cy = rng.uniform(0, W), rng.uniform(0, H)
        idx = rng.choice(n, n_hot, replace=False)
        x[idx] = np.clip(np.rint(rng.normal(cx, hot_sigma, n_hot)), 0, W - 1).astype(np.int64)
        y[idx] = np.clip(np.rint(rng.normal(cy, hot_sigma, n_hot)), 0, H - 1).astype(np.int64)
    p = (rng.random(n) < 0.5).astype(np.uint8)
    return x.astype(np.int16), y.astype(np.int16), t, p


def make_batch(cfg: int, B: int, H: int, W: int, n_lo: float, n_hi: float, first_sample: int = 0,
               fixed_n: int | None = None):
    """B windows, sizes ~ LogUniform[n_lo, n_hi]; seed = 1234 + 1000*cfg + sample index."""
    xs, ys, ts, ps, offs = [], [], [], [], [0]
    for i in range(B):
        rng = np.random.default_rng(1234 + 1000 * cfg + first_sample + i)
        n = fixed_n if fixed_n is not None else int(np.exp(rng.uniform(np.log(n_lo), np.log(n_hi))))
        x, y, t, p = make_window(rng, n, H, W)
        xs.append(x), ys.append(y), ts.append(t), ps.append(p)
        offs.append(offs[-1] + n)
    return (np.concatenate(xs), np.concatenate(ys), np.concatenate(ts), np.concatenate(ps),
            np.asarray(offs, dtype=np.int64))


def gen1_batch(B: int, cfg: int = 2, first_sample: int = 0):
    """Gen1-rate windows: N ~ LogUniform[2e4, 2e5] at 240x304."""
    return make_batch(cfg, B, GEN1[0], GEN1[1], 2e4, 2e5, first_sample)


def mpx_batch(B: int, cfg: int = 3, first_sample: int = 0):
    """1Mpx-rate windows: N ~ LogUniform[5e5, 5e6] at 720x1280."""
    return make_batch(cfg, B, MPX[0], MPX[1], 5e5, 5e6, first_sample)
