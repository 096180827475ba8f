"""Oracle (TEST INFRASTRUCTURE): PSEE ``.dat`` Event2D records -> event windows, numpy restatement.

Follows the reference's loader stack for one labelled timestamp:
  * record format and decode: ``yolox/utils/psee_loader/io/dat_events_tools.py:24, 40-51``
    (8 bytes: ``t:u4`` then ``_:i4`` with ``x = _ & 16383``, ``y = (_ >> 14) & 16383``,
    ``p = (_ >> 28) & 1``);
  * ``PSEELoader.seek_time`` (``psee_loader.py:196-238``) including its quirk: while more than
    ``term_criterion`` (100000) events remain the bisection probes ``t[middle]`` and, on an exact hit,
    returns with the file cursor one event PAST ``middle`` (the probe read advanced it);
  * ``PSEELoader.load_delta_t`` (``psee_loader.py:128-171``): events from the cursor up to the first one
    with ``t >= current_time + delta_t``;
  * ``GEN1Dataset.search_events`` (``gen1.py:217-236``): window ``[t + w0, t + w1)``, moved back by its own
    length while it is empty, at most ``num_slice + 1`` times (the ``zero_trigger`` loop).
Pinned by ``tests/golden/psee.npz`` (the reference's own PSEELoader run on synthetic ``.dat`` files written
with the reference's writer, ``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import numpy as np

TERM_CRITERION = 100000   # psee_loader.py:196 default
EV_DTYPE = np.dtype([("t", "<u4"), ("_", "<i4")])


def pack_records(x, y, t, p) -> np.ndarray:
    """``write_event_buffer`` (dat_events_tools.py:210-233): -> structured array of 8-byte records."""
    rec = np.empty(len(t), dtype=EV_DTYPE)
    rec["t"] = np.asarray(t).astype(np.uint32)
    rec["_"] = (np.asarray(x).astype(np.int32) + (np.asarray(y).astype(np.int32) << 14)
                + ((np.asarray(p) == 1).astype(np.int32) << 28))
    return rec


def decode(rec: np.ndarray):
    """(x:u2, y:u2, t:u4, p:u1) of a record array (dat_events_tools.py:46-51, psee_loader.py:43-48)."""
    w = rec["_"]
    x = np.bitwise_and(w, 16383).astype(np.uint16)
    y = np.right_shift(np.bitwise_and(w, 268419072), 14).astype(np.uint16)
    p = np.right_shift(np.bitwise_and(w, 268435456), 28).astype(np.uint8)
    return x, y, rec["t"], p


def seek_time(t: np.ndarray, final_time: int):
    """-> (cursor, current_time, done) after ``PSEELoader.seek_time(final_time)``."""
    n = len(t)
    total = int(t[-1]) if n else 0          # total_time(): timestamp of the last event
    if final_time > total:
        return n, total + 1, True
    if final_time <= 0:
        return 0, 0, False
    low, high = 0, n
    while high - low > TERM_CRITERION:
        middle = (low + high) // 2
        mid = int(t[middle])
        if mid > final_time:
            high = middle
        elif mid < final_time:
            low = middle + 1
        else:                                # exact hit: the probe read left the cursor at middle + 1
            return middle + 1, final_time, middle + 1 >= n
    pos = low + int(np.searchsorted(t[low:high], final_time))
    return pos, final_time, pos >= n


def load_delta_t(t: np.ndarray, cursor: int, current_time: int, done: bool, delta_t: int):
    """-> (lo, hi) record range returned by ``PSEELoader.load_delta_t(delta_t)``."""
    n = len(t)
    if done or cursor >= n:
        return cursor, cursor
    final_time = current_time + delta_t
    hi = cursor + int(np.searchsorted(t[cursor:], final_time))
    return cursor, hi


def search_events(t: np.ndarray, timestamp: int, window=(-50000, 0), num_slice: int = 1):
    """``GEN1Dataset.search_events`` for the 'fix_t' policy -> (lo, hi)."""
    delta = window[1] - window[0]
    cur = int(timestamp) + window[0]
    zero_trigger = 0
    while True:
        cursor, now, done = seek_time(t, cur)
        lo, hi = load_delta_t(t, cursor, now, done, delta)
        if hi - lo > 0 or zero_trigger > num_slice:
            return lo, hi
        zero_trigger += 1
        cur -= delta


def windows(rec: np.ndarray, timestamps, window=(-50000, 0), num_slice: int = 1) -> np.ndarray:
    t = rec["t"]
    return np.array([search_events(t, int(ts), window, num_slice) for ts in timestamps], dtype=np.int64).reshape(-1, 2)


def micro_sum_windows(rec: np.ndarray, ranges: np.ndarray, H: int, W: int, Tm: int) -> np.ndarray:
    """``agrregate(events, 'micro_sum')`` of every record range -> ``[B, Tm, 2, H, W]`` (gen1.py:355-360 on
    the decoded events; an empty range is the reference's ``None`` / empty case: zeros)."""
    from . import binning
    out = np.zeros((len(ranges), Tm, 2, H, W))
    for b, (lo, hi) in enumerate(ranges):
        if hi > lo:
            x, y, t, p = decode(rec[lo:hi])
            out[b] = binning.micro_sum(x, y, t, p, H, W, Tm)
    return out
