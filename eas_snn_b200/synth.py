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


def dat_stream(seed: int, n: int, H: int, W: int, span_us: int, t_start: int = 1000, gap=None):
    """A time-sorted synthetic recording for the PSEE ``.dat`` path: (x, y, t, p) with ``t`` in
    ``[t_start, t_start + span_us)``; ``gap = (a, b)`` removes every event with ``a <= t < b``.
    Regenerated from the seed by the golden script and by the tests (the fixture stores outputs only)."""
    rng = np.random.default_rng(seed)
    t = np.sort(rng.integers(t_start, t_start + span_us, int(n), dtype=np.int64))
    x = rng.integers(0, W, int(n), dtype=np.int64)
    y = rng.integers(0, H, int(n), dtype=np.int64)
    p = (rng.random(int(n)) < 0.5).astype(np.uint8)
    if gap is not None:
        keep = (t < gap[0]) | (t >= gap[1])
        x, y, t, p = x[keep], y[keep], t[keep], p[keep]
    return x.astype(np.int16), y.astype(np.int16), t, p


# name -> (stream kwargs, window (us), num_slice, Tm): shared by tests/golden/make_golden.py and the tests
DAT_CASES = {
    "small_24x32": (dict(seed=11, n=30_000, H=24, W=32, span_us=2_000_000), (-50_000, 0), 1, 4),
    "gap_backoff": (dict(seed=12, n=40_000, H=24, W=32, span_us=2_000_000, gap=(600_000, 1_000_000)), (-50_000, 0), 3, 4),
    "dense_bisect": (dict(seed=13, n=260_000, H=24, W=32, span_us=1_000_000), (-20_000, 0), 1, 5),
    "long_window": (dict(seed=14, n=50_000, H=40, W=48, span_us=1_500_000, t_start=0), (-200_000, 0), 2, 4),
}


def dat_label_times(name: str, t: np.ndarray, window, seed: int = 0):
    """Label timestamps that exercise every branch of the reference's window search: before the first
    event, inside, exactly on event timestamps, inside a hole, past the end, and (for streams with more
    than 100000 events) exactly on the bisection probes of PSEELoader.seek_time."""
    rng = np.random.default_rng(1000 + seed)
    n = len(t)
    ts = [int(t[0]) - 10, int(t[0]) + 5, int(t[0]) - window[0], int(t[n // 3]) - window[0], int(t[n // 2]),
          int(t[-1]), int(t[-1]) - window[0], int(t[-1]) - window[0] + 1, int(t[-1]) + 10 * (window[1] - window[0])]
    ts += [int(v) for v in rng.integers(int(t[0]), int(t[-1]), 12)]
    if name == "gap_backoff":
        ts += [650_000, 700_000, 790_000, 800_000, 860_000, 990_000, 1_000_000, 1_049_000]
    lo, hi = 0, n
    while hi - lo > 100000:                       # the probes of the reference's bisection for a late target
        mid = (lo + hi) // 2
        ts += [int(t[mid]) - window[0], int(t[mid]) - window[0] + 1]
        lo = mid + 1
    return np.asarray(ts, dtype=np.int64)
