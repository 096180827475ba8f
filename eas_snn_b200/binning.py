"""Event binning on the GPU: ``(x, y, t, p)`` windows -> ``[B, Tm, 2, H, W]`` int32 histograms.

Drop-in for the reference's CPU path ``GEN1Dataset.slice_events`` + ``agrregate('micro_sum')``
(``yolox/data/datasets/gen1.py:313-360``), called through ``eas_bin_events`` of the C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

STRATEGY = {"auto": 0, "reds": 1, "tiles": 2}


def bin_events(x: torch.Tensor, y: torch.Tensor, t: torch.Tensor, p: torch.Tensor, offsets: torch.Tensor,
               H: int, W: int, Tm: int, strategy: str = "auto", out: torch.Tensor | None = None,
               dtype: torch.dtype = torch.int32) -> torch.Tensor:
    """Histogram B time-sorted windows.

    x, y : int16 ``[N]``; t : int64 ``[N]`` (sorted inside each window); p : uint8/bool ``[N]``;
    offsets : int64 ``[B+1]`` with ``offsets[0] == 0`` and ``offsets[B] == N``.  All CUDA tensors.
    Returns ``[B, Tm, 2, H, W]`` counts, int32 (default) or float32 (``dtype=torch.float32``: exact
    below 2^24, the dtype the reference casts to on the device and the sampler consumes).
    """
    _lib.require_cuda(x, y, t, p, offsets)
    if p.dtype == torch.bool:
        p = p.view(torch.uint8)
    if x.dtype != torch.int16 or y.dtype != torch.int16 or t.dtype != torch.int64 or p.dtype != torch.uint8:
        raise TypeError("bin_events expects x,y int16, t int64, p uint8/bool (events_struct, util.py:119-121)")
    if offsets.dtype != torch.int64 or offsets.dim() != 1 or offsets.numel() < 1:
        raise TypeError("offsets must be int64 [B+1]")
    n = x.numel()
    if not (y.numel() == n and t.numel() == n and p.numel() == n):
        raise ValueError("x, y, t, p must have the same length")
    x, y, t, p, offsets = (a.contiguous() for a in (x, y, t, p, offsets))
    B = offsets.numel() - 1
    if out is None:
        if dtype not in (torch.int32, torch.float32):
            raise TypeError("histogram dtype must be int32 or float32")
        out = torch.empty((B, Tm, 2, H, W), dtype=dtype, device=x.device)
    elif (out.shape != (B, Tm, 2, H, W) or out.dtype not in (torch.int32, torch.float32)
          or not out.is_contiguous()):
        raise ValueError("out must be a contiguous int32/float32 [B, Tm, 2, H, W] tensor")
    L = _lib.lib()
    ws_bytes = L.eas_bin_events_ws_bytes(B, Tm)
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = L.eas_bin_events_ex(_lib.ptr(x), _lib.ptr(y), _lib.ptr(t), _lib.ptr(p), _lib.ptr(offsets),
                                 B, n, H, W, Tm, _lib.ptr(out), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(),
                                 STRATEGY[strategy],
                                 _lib.EAS_F32 if out.dtype == torch.float32 else _lib.EAS_I32)
    _lib.check(rc, "eas_bin_events")
    return out


class HostEventBatch:
    """Pinned host staging for one batch of windows (the host side of the e2e path)."""

    def __init__(self, x: np.ndarray, y: np.ndarray, t: np.ndarray, p: np.ndarray, offsets: np.ndarray):
        self.n = int(x.shape[0])
        self.B = int(offsets.shape[0]) - 1
        self.x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.int16)).pin_memory()
        self.y = torch.from_numpy(np.ascontiguousarray(y, dtype=np.int16)).pin_memory()
        self.t = torch.from_numpy(np.ascontiguousarray(t, dtype=np.int64)).pin_memory()
        self.p = torch.from_numpy(np.ascontiguousarray(p).astype(np.uint8)).pin_memory()
        self.offsets = torch.from_numpy(np.ascontiguousarray(offsets, dtype=np.int64)).pin_memory()

    @property
    def nbytes(self) -> int:
        return self.n * 13 + (self.B + 1) * 8

    def to_device(self, device):
        return tuple(a.to(device, non_blocking=True) for a in (self.x, self.y, self.t, self.p, self.offsets))


def rvt_event_sum(repr_u8: torch.Tensor, n_bins: int = 10) -> torch.Tensor:
    """RVT stacked histograms ``uint8 [n, 2*n_bins, H, W]`` (channel = polarity*n_bins + bin) -> per-polarity
    counts ``float32 [n, 2, H, W]``: the reference's ``'event_sum'`` reduction
    (``yolox/data/datasets/rvt_gen4.py:120-122``), on the GPU."""
    _lib.require_cuda(repr_u8)
    if repr_u8.dtype != torch.uint8 or repr_u8.dim() != 4 or repr_u8.shape[1] != 2 * n_bins:
        raise TypeError("expected uint8 [n, 2*n_bins, H, W]")
    repr_u8 = repr_u8.contiguous()
    n, _, H, W = repr_u8.shape
    out = torch.empty((n, 2, H, W), dtype=torch.float32, device=repr_u8.device)
    with torch.cuda.device(repr_u8.device):
        rc = _lib.lib().eas_rvt_event_sum(_lib.ptr(repr_u8), n, n_bins, H, W, _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "eas_rvt_event_sum")
    return out


def voxel_grid(x, y, t, p, offsets, H: int, W: int, n_bins: int = 10, polarity: str = "reference") -> torch.Tensor:
    """Voxel grids with bilinear interpolation in time (``to_voxel_grid_numpy``, ``yolox/utils/event_reps.py:30-89``) of
    B event windows (same SoA layout as :func:`bin_events`) -> ``float32 [B, n_bins, 1, H, W]`` (the reference returns
    ``(n_bins, 1, H, W)`` per window).  ``polarity``: ``"reference"`` reproduces what the reference computes on its own
    ``events_struct`` dtype, where ``p`` is bool and ``pols[pols == 0] = -1`` stores True -- every event counts +1
    (event_reps.py:62-63); ``"signed"`` is the +1 / -1 weighting its comment intends (and what it computes when ``p``
    is a signed integer field)."""
    if polarity not in ("reference", "signed"):
        raise ValueError("polarity must be 'reference' or 'signed'")
    _lib.require_cuda(x, y, t, p, offsets)
    if x.dtype != torch.int16 or y.dtype != torch.int16 or t.dtype != torch.int64 or p.dtype not in (torch.uint8, torch.bool) \
            or offsets.dtype != torch.int64:
        raise TypeError("expected x,y int16, t int64, p uint8/bool, offsets int64")
    B = offsets.numel() - 1
    out = torch.empty((B, n_bins, 1, H, W), dtype=torch.float32, device=x.device)
    p8 = p.view(torch.uint8) if p.dtype == torch.bool else p
    with torch.cuda.device(x.device):
        rc = _lib.lib().eas_voxel_grid(_lib.ptr(x.contiguous()), _lib.ptr(y.contiguous()), _lib.ptr(t.contiguous()),
                                       _lib.ptr(p8.contiguous()), _lib.ptr(offsets.contiguous()), B, x.numel(), H, W, n_bins,
                                       int(polarity == "signed"), _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "eas_voxel_grid")
    return out


_TAPS: dict = {}


def _linear_taps(n_out: int, n_in: int, device):
    """cv2.INTER_LINEAR taps of one axis (half-pixel centres in float32, taps clamped to the image): source index
    and float32 weight of the second tap, cached per (n_out, n_in, device)."""
    key = (n_out, n_in, str(device))
    if key not in _TAPS:
        scale = 1.0 / (float(n_out) / float(n_in))
        f = ((np.arange(n_out, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
        i0 = np.floor(f).astype(np.int64)
        f = (f - i0.astype(np.float32)).astype(np.float32)
        f[i0 < 0] = 0.0
        i0[i0 < 0] = 0
        f[i0 >= n_in - 1] = 0.0
        i0[i0 >= n_in - 1] = n_in - 1
        _TAPS[key] = (torch.from_numpy(i0.astype(np.int32)).to(device), torch.from_numpy(f).to(device))
    return _TAPS[key]


def letterbox_frames(frames: torch.Tensor, size, letterbox: bool = True, center: bool = False) -> torch.Tensor:
    """Letterbox + bilinear resize of micro-frames on the GPU: what ``GEN1Dataset.get_random_data(random=False)``
    (``yolox/data/datasets/gen1.py:433-483``) does on the host with ``cv2.resize(INTER_LINEAR)`` -- every
    ``[ih, iw]`` plane of ``frames [..., ih, iw]`` (fp32 or int32 counts) scaled by ``min(w/iw, h/ih)``, pasted
    top-left (``center``: centred) into a zero ``[h, w]`` canvas; ``letterbox=False`` stretches to ``(h, w)``.
    Counts become fractional, so this sits outside the bit-exact histogram contract (optional stage)."""
    _lib.require_cuda(frames)
    h, w = int(size[0]), int(size[1])
    ih, iw = frames.shape[-2:]
    if frames.dtype not in (torch.float32, torch.int32):
        frames = frames.float()
    frames = frames.contiguous()
    if letterbox:
        scale = min(w / iw, h / ih)
        nw, nh = int(iw * scale), int(ih * scale)
        dy, dx = ((h - nh) // 2, (w - nw) // 2) if center else (0, 0)
    else:
        nh, nw, dy, dx = h, w, 0, 0
    if w % 4 != 0:
        raise ValueError("letterbox_frames: the output width must be a multiple of 4")
    x0, fx = _linear_taps(nw, iw, frames.device)
    y0, fy = _linear_taps(nh, ih, frames.device)
    out = torch.empty(frames.shape[:-2] + (h, w), dtype=torch.float32, device=frames.device)
    n_planes = frames.numel() // (ih * iw)
    in_dtype = _lib.EAS_I32 if frames.dtype == torch.int32 else _lib.EAS_F32
    with torch.cuda.device(frames.device):
        rc = _lib.lib().eas_letterbox_bilinear(_lib.ptr(frames), in_dtype, n_planes, ih, iw, _lib.ptr(x0), _lib.ptr(fx),
                                               _lib.ptr(y0), _lib.ptr(fy), nh, nw, dy, dx, _lib.ptr(out), h, w,
                                               _lib.stream_ptr())
    _lib.check(rc, "eas_letterbox_bilinear")
    return out
