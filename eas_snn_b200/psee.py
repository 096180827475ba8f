"""PSEE ``.dat`` recordings on the GPU: raw records in, event windows / histograms out.

Replaces, for the hot path, the reference's CPU loader stack
``PSEELoader`` (``yolox/utils/psee_loader/io/psee_loader.py:30-238``), its record decode
(``dat_events_tools.py:24, 40-51``) and ``GEN1Dataset.search_events`` (``gen1.py:217-236``): the 8-byte
Event2D records of a whole recording stay resident in HBM, window search (``seek_time`` +
``load_delta_t`` + the empty-window back-off) is one kernel for a batch of label timestamps, and binning
reads the packed records directly (8 B/event instead of the 13 B of the decoded struct).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .binning import STRATEGY

RECORD_DTYPE = np.dtype([("t", "<u4"), ("_", "<i4")])   # dat_events_tools.py:24 (Event2D)


def pack_records(x, y, t, p) -> np.ndarray:
    """Host helper: (x, y, t, p) arrays -> ``[n, 2]`` int32 array of raw records (column 0 = t as the bit
    pattern of a uint32, column 1 = ``x | y << 14 | p << 28``), the layout of a ``_td.dat`` payload."""
    t = np.asarray(t)
    rec = np.empty((len(t), 2), dtype=np.int32)
    rec[:, 0] = t.astype(np.uint32).view(np.int32)
    rec[:, 1] = (np.asarray(x).astype(np.int32) | (np.asarray(y).astype(np.int32) << 14)
                 | ((np.asarray(p) != 0).astype(np.int32) << 28))
    return rec


def read_dat(path: str):
    """Parse a ``_td.dat`` file: text header lines starting with ``%``, then (when a header exists) one
    byte event type and one byte event size, then the records.  Returns ``(records [n, 2] int32, (H, W))``."""
    height = width = None
    with open(path, "rb") as f:
        n_comment = 0
        while True:
            pos = f.tell()
            line = f.readline()
            if line[:2] != b"% ":
                f.seek(pos)
                break
            n_comment += 1
            words = line.split()
            if len(words) > 2 and words[1] == b"Height":
                height = int(words[2])
            if len(words) > 2 and words[1] == b"Width":
                width = int(words[2])
        ev_size = 8
        if n_comment > 0:
            ev_type, ev_size = (int(v) for v in f.read(2))
            if ev_type != 0:
                raise ValueError("%s: event type %d is not Event2D" % (path, ev_type))
        if ev_size != 8:
            raise ValueError("%s: event size %d, expected 8 bytes" % (path, ev_size))
        payload = np.fromfile(f, dtype=np.int32)
    if payload.size % 2:
        raise ValueError("%s: truncated record" % path)
    return payload.reshape(-1, 2), (height, width)


def _records(records: torch.Tensor) -> torch.Tensor:
    _lib.require_cuda(records)
    if records.dtype != torch.int32 or records.dim() != 2 or records.shape[1] != 2:
        raise TypeError("records must be an int32 [n, 2] CUDA tensor of raw .dat Event2D records")
    return records.contiguous()


def dat_windows(records: torch.Tensor, t_label: torch.Tensor, window=(-50000, 0), max_backoff: int = 1) -> torch.Tensor:
    """Record range ``[B, 2]`` (int64) of the event window of every label timestamp (``search_events``)."""
    records = _records(records)
    _lib.require_cuda(t_label)
    if t_label.dtype != torch.int64 or t_label.dim() != 1:
        raise TypeError("t_label must be int64 [B]")
    if not window[1] > window[0]:
        raise ValueError("window must be (start, end) with end > start")
    t_label = t_label.contiguous()
    ranges = torch.empty((t_label.numel(), 2), dtype=torch.int64, device=records.device)
    with torch.cuda.device(records.device):
        rc = _lib.lib().eas_dat_windows(_lib.ptr(records), records.shape[0], _lib.ptr(t_label), t_label.numel(),
                                        int(window[0]), int(window[1]), int(max_backoff), _lib.ptr(ranges),
                                        _lib.stream_ptr())
    _lib.check(rc, "eas_dat_windows")
    return ranges


def bin_dat(records: torch.Tensor, ranges: torch.Tensor, H: int, W: int, Tm: int, strategy: str = "auto",
            out=None, dtype: torch.dtype = torch.int32):
    """Histogram record ranges: ``[B, Tm, 2, H, W]`` int32 / float32 counts (or the compact byte form,
    ``dtype=torch.uint8`` -> :class:`eas_snn_b200.binning.CompactHist`), bit-identical to decoding the
    records and calling :func:`eas_snn_b200.bin_events` (``agrregate('micro_sum')``, gen1.py:313-360)."""
    records = _records(records)
    _lib.require_cuda(ranges)
    if ranges.dtype != torch.int64 or ranges.dim() != 2 or ranges.shape[1] != 2:
        raise TypeError("ranges must be int64 [B, 2]")
    ranges = ranges.contiguous()
    B = ranges.shape[0]
    from .binning import CompactHist, _hist_out, poll_compact
    out, buf, eas_dtype = _hist_out(B, Tm, H, W, out, dtype, records.device)
    L = _lib.lib()
    ws_bytes = L.eas_bin_dat_ws_bytes(B, Tm)
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=records.device)
    with torch.cuda.device(records.device):
        rc = L.eas_bin_dat(_lib.ptr(records), records.shape[0], _lib.ptr(ranges), B, H, W, Tm, _lib.ptr(buf),
                           _lib.ptr(ws), ws_bytes, _lib.stream_ptr(), STRATEGY[strategy], eas_dtype)
    _lib.check(rc, "eas_bin_dat")
    if isinstance(out, CompactHist) and B > 0:
        poll_compact()
        out._queue_check()
    return out


class DatRecording:
    """A ``_td.dat`` recording resident in GPU memory, with the slice of ``PSEELoader`` the dataset uses."""

    def __init__(self, records: torch.Tensor, height: int | None = None, width: int | None = None):
        self.records = _records(records)
        self.height, self.width = height, width

    @classmethod
    def from_file(cls, path: str, device) -> "DatRecording":
        rec, (h, w) = read_dat(path)
        return cls(torch.from_numpy(rec).to(device), h, w)

    @classmethod
    def from_events(cls, x, y, t, p, device, height=None, width=None) -> "DatRecording":
        return cls(torch.from_numpy(pack_records(x, y, t, p)).to(device), height, width)

    def event_count(self) -> int:
        return int(self.records.shape[0])

    def windows(self, t_label, window=(-50000, 0), num_slice: int = 1) -> torch.Tensor:
        """``GEN1Dataset.search_events`` for a batch of label timestamps -> record ranges ``[B, 2]``."""
        if not torch.is_tensor(t_label):
            t_label = torch.as_tensor(np.asarray(t_label, dtype=np.int64))
        return dat_windows(self.records, t_label.to(self.records.device), window, num_slice)

    def histograms(self, ranges: torch.Tensor, H: int, W: int, Tm: int, **kw) -> torch.Tensor:
        return bin_dat(self.records, ranges, H, W, Tm, **kw)
