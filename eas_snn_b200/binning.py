"""Event binning on the GPU: ``(x, y, t, p)`` windows -> ``[B, Tm, 2, H, W]`` int32 histograms.

Drop-in for the reference's CPU path ``GEN1Dataset.slice_events`` + ``agrregate('micro_sum')``
(``yolox/data/datasets/gen1.py:313-360``), called through ``eas_bin_events`` of the C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

STRATEGY = {"auto": 0, "reds": 1, "tiles": 2, "tiles_pair": 3, "tiles_planes": 4}
SAT_CAP = 4096          # EAS_HIST_U8_SAT_CAP (include/eas_b200.h)


class CompactHist:
    """The byte histogram of ``bin_events(..., dtype=torch.uint8)`` (``include/eas_b200.h``, EAS_U8): one device buffer
    holding ``min(count, 255)`` per bin followed by the exact ``{bin, count}`` list of the bins that reached 255 -- the
    information of the int32 histogram in a quarter of the bytes.  The sampler kernels read it directly
    (``AdaptiveRSNNEmbedding.forward(CompactHist)``); :meth:`dense` gives the ``[B, Tm, 2, H, W]`` tensor back.

    Only when more than ``SAT_CAP`` bins saturate in one call is information lost (``lost`` flag in the buffer).
    That is reported without stalling the stream: the binning kernel sets a pinned host int (registered with the library
    once per process) at the moment it loses a count, and :func:`poll_compact` (run by the next binning call, or by hand)
    raises ``OverflowError`` once it is set; :meth:`check` waits and checks this buffer now."""

    def __init__(self, buf: torch.Tensor, shape):
        self.buf, self.shape = buf, tuple(int(v) for v in shape)
        self.nbins = int(np.prod(self.shape))
        self._tail_off = (self.nbins + 255) // 256 * 256

    @staticmethod
    def nbytes_for(shape) -> int:
        nbins = int(np.prod(shape))
        return (nbins + 255) // 256 * 256 + 16 + SAT_CAP * 8

    @classmethod
    def empty(cls, shape, device) -> "CompactHist":
        return cls(torch.empty(cls.nbytes_for(shape), dtype=torch.uint8, device=device), shape)

    @property
    def device(self):
        return self.buf.device

    @property
    def counts(self) -> torch.Tensor:
        """uint8 ``[B, Tm, 2, H, W]`` view: the counts, 255 where a bin saturated."""
        return self.buf[:self.nbins].view(self.shape)

    @property
    def tail(self) -> torch.Tensor:
        """int32 view of the list header: ``[n_saturated, lost, 0, 0]``."""
        return self.buf[self._tail_off:self._tail_off + 16].view(torch.int32)

    def saturated(self):
        """(bin indices, exact counts) of the saturated bins (synchronises)."""
        n = min(int(self.tail[0]), SAT_CAP)
        ent = self.buf[self._tail_off + 16:self._tail_off + 16 + 8 * n].view(torch.int32).view(n, 2)
        return ent[:, 0].to(torch.int64) & 0xFFFFFFFF, ent[:, 1].clone()

    def dense(self, dtype: torch.dtype = torch.int32) -> torch.Tensor:
        """The ``[B, Tm, 2, H, W]`` int32 / float32 histogram (``eas_hist_u8_expand``)."""
        if dtype not in (torch.int32, torch.float32):
            raise TypeError("dense histogram dtype must be int32 or float32")
        out = torch.empty(self.shape, dtype=dtype, device=self.buf.device)
        B, Tm, _, H, W = self.shape
        with torch.cuda.device(self.buf.device):
            rc = _lib.lib().eas_hist_u8_expand(_lib.ptr(self.buf), B, Tm, H, W, _lib.ptr(out),
                                               _lib.EAS_F32 if dtype == torch.float32 else _lib.EAS_I32,
                                               _lib.stream_ptr())
        _lib.check(rc, "eas_hist_u8_expand")
        return out

    def check(self):
        """Wait for the binning call and raise if it lost counts."""
        if int(self.tail[1]) != 0:
            raise OverflowError(_LOST_MSG)
        return self

    def _queue_check(self):
        """Nothing to launch: the binning kernel itself sets the registered pinned host int (``_sticky_flag``) at the
        moment it loses a count; :func:`poll_compact` reads it."""
        _sticky_flag()


_LOST_MSG = ("compact histogram: more than %d bins collected >= 255 events in one call, counts were lost; "
             "bin with dtype=torch.int32 / float32 (hist_dtype='dense')" % SAT_CAP)
_STICKY: list = []       # one pinned int32: set by the device when a compact binning call lost counts


def _sticky_flag() -> torch.Tensor:
    """The process-wide pinned int32 the compact binning kernels set when they lose a count (registered with the
    library once: ``eas_hist_u8_set_sticky``; pinned host memory is device-accessible under unified addressing)."""
    if not _STICKY:
        flag = torch.zeros(1, dtype=torch.int32).pin_memory()
        _lib.check(_lib.lib().eas_hist_u8_set_sticky(C.c_void_p(flag.data_ptr())), "eas_hist_u8_set_sticky")
        _STICKY.append(flag)
    return _STICKY[0]


def poll_compact(wait: bool = False):
    """Raise ``OverflowError`` (once) if a compact binning call that has run so far lost counts; ``wait=True``
    synchronises the device first, so that every call issued so far is covered."""
    if wait:
        torch.cuda.synchronize()
    if _STICKY and int(_STICKY[0][0]) != 0:
        _STICKY[0][0] = 0
        raise OverflowError(_LOST_MSG)


def _hist_out(B, Tm, H, W, out, dtype, device):
    """Allocate / validate the output of a binning call; returns (object handed back, buffer, EAS dtype)."""
    shape = (B, Tm, 2, H, W)
    if out is None:
        if dtype == torch.uint8:
            out = CompactHist.empty(shape, device)
        elif dtype in (torch.int32, torch.float32):
            out = torch.empty(shape, dtype=dtype, device=device)
        else:
            raise TypeError("histogram dtype must be int32, float32 or uint8 (compact)")
    if isinstance(out, CompactHist):
        if out.shape != shape or out.buf.numel() < CompactHist.nbytes_for(shape):
            raise ValueError("out: CompactHist of another shape")
        _sticky_flag()                       # (registered with the library before the first compact call)
        return out, out.buf, _lib.EAS_U8
    if out.shape != shape or out.dtype not in (torch.int32, torch.float32) or not out.is_contiguous():
        raise ValueError("out must be a contiguous int32/float32 [B, Tm, 2, H, W] tensor or a CompactHist")
    return out, out, _lib.EAS_F32 if out.dtype == torch.float32 else _lib.EAS_I32


def bin_events(x: torch.Tensor, y: torch.Tensor, t: torch.Tensor, p: torch.Tensor, offsets: torch.Tensor,
               H: int, W: int, Tm: int, strategy: str = "auto", out=None,
               dtype: torch.dtype = torch.int32):
    """Histogram B time-sorted windows.

    x, y : int16 ``[N]``; t : int64 ``[N]`` (sorted inside each window); p : uint8/bool ``[N]``;
    offsets : int64 ``[B+1]`` with ``offsets[0] == 0`` and ``offsets[B] == N``.  All CUDA tensors.
    Returns ``[B, Tm, 2, H, W]`` counts, int32 (default) or float32 (``dtype=torch.float32``: exact
    below 2^24, the dtype the reference casts to on the device and the sampler consumes), or -- with
    ``dtype=torch.uint8`` -- a :class:`CompactHist` (bytes + exact saturation list; frames that fit the shared-memory
    tiles kernel only).
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
    out, buf, eas_dtype = _hist_out(B, Tm, H, W, out, dtype, x.device)
    L = _lib.lib()
    ws_bytes = L.eas_bin_events_ws_bytes(B, Tm)
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = L.eas_bin_events_ex(_lib.ptr(x), _lib.ptr(y), _lib.ptr(t), _lib.ptr(p), _lib.ptr(offsets),
                                 B, n, H, W, Tm, _lib.ptr(buf), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(),
                                 STRATEGY[strategy], eas_dtype)
    _lib.check(rc, "eas_bin_events")
    if isinstance(out, CompactHist) and B > 0:
        poll_compact()
        out._queue_check()
    return out


def compact_fits(H: int, W: int) -> int:
    """Whether the byte histogram can be written for this frame size: the shared-memory tiles kernel holds a row slab
    of 16-bit counters per CTA (<= 8 slabs of <= 72 KB; bin_events.cu ``slab_geo``).  Returns the number of row slabs
    (work items per window = Tm * 2 * slabs), 0 when the frame does not fit."""
    slab = 72 * 1024
    n_slabs = min(-(-(H * W * 2) // slab), H)
    rows = -(-H // n_slabs)
    n_slabs = -(-H // rows)
    smem = ((rows * W + 1) // 2 + 3) // 4 * 4 * 4
    return n_slabs if (smem <= slab + 4096 and n_slabs <= 8) else 0


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
