"""``AdaptiveRSNNEmbedding``: the EAS adaptive event sampler on sm_100a kernels.

Drop-in for ``yolox.models.embedding.AdaptiveRSNNEmbedding`` (``yolox/models/embedding.py:79-226``):
same constructor, same parameter / state-dict layout (``gate_conv.{0,2}.*``, ``input_conv.{0,2}.*``),
same ``forward`` contract (4-D passthrough, 5-D ``[B,Tm,2,H,W]``, 6-D ``[B,Tl,Tm,2,H,W]`` ->
``[Ts, B*Tl, 2, H, W]`` fp32).  New front door: :meth:`forward_events` takes the raw
``(x, y, t, p)`` stream and runs binning + sampling without the dense tensor ever leaving the GPU
or being converted to floating point.
"""
from __future__ import annotations

import copy
import ctypes as C
import os

import torch
import torch.nn as nn

from . import _lib
from .binning import CompactHist, bin_events, compact_fits


def _cfg(B, H, W, Tm, mod, in_dtype):
    return _lib.SamplerCfg(B=B, H=H, W=W, Tm=Tm, Ts=mod.Ts, ksize=mod.kernel_size, depth=mod.depth,
                           readout=_lib.READOUT[mod.readout], hard_reset=0 if mod.vreset is None else 1,
                           vreset=0.0 if mod.vreset is None else float(mod.vreset), thresh=float(mod.thresh),
                           spike_attach=int(bool(mod.spike_attach)), write_zero=int(bool(mod.write_zero)),
                           use_abs=int(bool(mod.abs)), in_dtype=in_dtype,
                           algo=_lib.SAMPLER_ALGO[getattr(mod, "algo", "auto")],
                           surr_alpha=float(getattr(mod.kwargs_spikes.get("spike_fn", None), "alpha", 1.0)))


def _pack_ptrs(tensors):
    """[in_w0, in_b0, in_w1, in_b1, gate_w0, gate_b0, gate_w1, gate_b1] (None allowed) -> struct."""
    s = _lib.SamplerPtrs()
    for name, t in zip(("in_w0", "in_b0", "in_w1", "in_b1", "gate_w0", "gate_b0", "gate_w1", "gate_b1"), tensors):
        setattr(s, name, None if t is None else t.data_ptr())
    return s


def _sampler_forward(events, mod, params, want_seq):
    """One ``eas_sampler_fwd`` call.  params: in_w0, in_b0, [in_w1, in_b1], gate_w0, gate_b0, [gate_w1, gate_b1].
    Returns (out [Ts,B,2,H,W], v_seq, gate_seq ([Tm,B,2,H,W] each, sampler step order, or None), cfg, plist)."""
    L = _lib.lib()
    B, Tm, Cc, H, W = events.shape
    if isinstance(events, CompactHist):
        in_dtype, ev_buf = _lib.EAS_U8, events.buf
    else:
        in_dtype, ev_buf = (_lib.EAS_I32 if events.dtype == torch.int32 else _lib.EAS_F32), events
    cfg = _cfg(B, H, W, Tm, mod, in_dtype)
    if mod.depth == 2:
        iw0, ib0, iw1, ib1, gw0, gb0, gw1, gb1 = params
    else:
        iw0, ib0, gw0, gb0 = params
        iw1 = ib1 = gw1 = gb1 = None
    plist = [iw0, ib0, iw1, ib1, gw0, gb0, gw1, gb1]
    plist = [None if q is None else q.detach().contiguous().float() for q in plist]
    wstruct = _pack_ptrs(plist)
    dev = events.device
    out = torch.empty((mod.Ts, B, 2, H, W), dtype=torch.float32, device=dev)
    v_seq = gate_seq = None
    if want_seq:
        v_seq = torch.empty((Tm, B, 2, H, W), dtype=torch.float32, device=dev)
        gate_seq = torch.empty_like(v_seq)
    ws_bytes = L.eas_sampler_fwd_ws_bytes(C.byref(cfg))
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = L.eas_sampler_fwd(C.byref(cfg), _lib.ptr(ev_buf), C.byref(wstruct), _lib.ptr(out),
                               _lib.ptr(v_seq), _lib.ptr(gate_seq), _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
    _lib.check(rc, "eas_sampler_fwd")
    return out, v_seq, gate_seq, cfg, plist


class _SamplerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, events, mod, need_grad, *params):
        # need_grad is decided by the caller: ctx.needs_input_grad is True for every Parameter even under
        # torch.no_grad(), and saving the per-step sequences costs 2 x [Tm, B, 2, H, W] of stores per forward
        out, v_seq, gate_seq, cfg, plist = _sampler_forward(events, mod, params, need_grad)
        if need_grad:
            ctx.mod, ctx.cfg, ctx.plist = mod, cfg, plist
            ctx.save_for_backward(events, v_seq, gate_seq)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        L = _lib.lib()
        events, v_seq, gate_seq = ctx.saved_tensors
        mod, cfg, plist = ctx.mod, ctx.cfg, ctx.plist
        dev = events.device
        grad_out = grad_out.contiguous().float()
        grads = [None if q is None else torch.empty_like(q) for q in plist]
        gstruct = _pack_ptrs(grads)
        wstruct = _pack_ptrs(plist)
        g_ev = None
        if ctx.needs_input_grad[0]:
            g_ev = torch.empty(events.shape, dtype=torch.float32, device=dev)
        ws_bytes = L.eas_sampler_bwd_ws_bytes(C.byref(cfg))
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = L.eas_sampler_bwd(C.byref(cfg), _lib.ptr(events), C.byref(wstruct), _lib.ptr(v_seq),
                                   _lib.ptr(gate_seq), _lib.ptr(grad_out), C.byref(gstruct), _lib.ptr(g_ev),
                                   _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
        _lib.check(rc, "eas_sampler_bwd")
        if mod.depth == 2:
            gp = grads
        else:
            gp = [grads[0], grads[1], grads[4], grads[5]]
        return (g_ev, None, None) + tuple(gp)


class AdaptiveRSNNEmbedding(nn.Module):
    """Same signature as the reference class (embedding.py:80-104)."""

    def __init__(self, kernel_size, in_channel=2, out_channel=2, Ts=1, split=False, spike_attach=False,
                 write_zero=False, abs=False, depth=1, readout="sum", **kwargs_spikes):
        super().__init__()
        if in_channel != 2 or out_channel != 2:
            # the reference's own state is zeros_like(events[0]) (embedding.py:159-160): out == in;
            # event polarity gives 2 channels (in_dim = 2, event_yolox_base.py:34)
            raise NotImplementedError("the sm_100a sampler is built for 2 polarity channels")
        if int(depth) not in (1, 2) or int(kernel_size) not in (3, 5, 7):
            raise NotImplementedError("sampler kernels are built for depth in {1,2}, kernel_size in {3,5,7}")
        if readout not in _lib.READOUT:
            raise NotImplementedError(readout)
        self.kernel_size = int(kernel_size)
        self.kwargs_spikes = kwargs_spikes
        self.Ts = int(Ts)
        self.abs = abs
        self.split = split
        self.readout = readout
        self.write_zero = write_zero
        self.nb_steps = kwargs_spikes["nb_steps"] if "Tm" not in kwargs_spikes else kwargs_spikes["Tm"]
        self.thresh = kwargs_spikes["thresh"]
        self.vreset = copy.deepcopy(kwargs_spikes["vreset"])
        fn = kwargs_spikes.get("spike_fn", None)
        if fn is not None and getattr(fn, "__name__", type(fn).__name__) != "Rectangle":
            raise NotImplementedError("the sampler kernel implements the Rectangle spike function "
                                      "(hard-wired by the reference, event_yolox_base.py:156)")
        self.depth = int(depth)
        self.gate_conv = self.build_conv(out_channel, out_channel * 2, self.kernel_size, depth=self.depth)
        self.input_conv = self.build_conv(in_channel, out_channel * 2, self.kernel_size, depth=self.depth)
        if self.split:  # created but never used by the reference forward (embedding.py:100-102)
            self.gate_conv_agg = nn.Conv2d(out_channel, out_channel * 2, self.kernel_size,
                                           padding=self.kernel_size // 2)
            self.input_conv_agg = nn.Conv2d(in_channel, out_channel * 2, self.kernel_size,
                                            padding=self.kernel_size // 2)
        self.spike_attach = spike_attach
        # forward kernel: "auto" (tensor cores for depth 2 / k 5, else the FP32-pipe kernel), "fp32", "tensor"
        self.algo = os.environ.get("EAS_SAMPLER_ALGO", "auto")
        self._init_weight()

    @staticmethod
    def build_conv(in_channel, out_channel, kernel_size, depth=1):
        convs = [nn.Conv2d(in_channel, out_channel, kernel_size, padding=kernel_size // 2)]
        for _ in range(depth - 1):
            convs.append(nn.ReLU(inplace=True))
            convs.append(nn.Conv2d(out_channel, out_channel, kernel_size, padding=kernel_size // 2))
        return nn.Sequential(*convs)

    def _init_weight(self):
        for m in self.input_conv.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.orthogonal_(m.weight, gain=nn.init.calculate_gain("relu"))
        for m in self.gate_conv.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, nonlinearity="sigmoid")

    def _params(self):
        ic = [m for m in self.input_conv if isinstance(m, nn.Conv2d)]
        gc = [m for m in self.gate_conv if isinstance(m, nn.Conv2d)]
        out = []
        for convs in (ic, gc):
            for m in convs:
                out += [m.weight, m.bias]
        return out

    def forward(self, events, record=False, v_record=False):
        if isinstance(events, CompactHist):   # the byte histogram of bin_events(dtype=torch.uint8): read as it is
            params = self._params()
            if record or v_record or self.algo == "tensor_split" or \
                    (torch.is_grad_enabled() and any(q.requires_grad for q in params)):
                return self.forward(events.dense(torch.float32), record, v_record)
            return _sampler_forward(events, self, params, False)[0]
        if events.dim() < 5:  # parameter-registering passthrough (embedding.py:144-146)
            events, _ = torch.broadcast_tensors(events, torch.zeros((self.Ts,) + events.shape,
                                                                    device=events.device))
            return events
        if events.dim() > 5:
            events = events.flatten(end_dim=-5)
        _lib.require_cuda(events)
        if events.shape[2] != 2:
            raise ValueError("expected [B, Tm, 2, H, W] micro-bin tensor")
        if events.dtype not in (torch.float32, torch.int32):
            events = events.float()
        events = events.contiguous()
        if record or v_record:
            return self._forward_with_record(events, record)
        params = self._params()
        need_grad = torch.is_grad_enabled() and (events.requires_grad or any(q.requires_grad for q in params))
        return _SamplerFn.apply(events, self, need_grad, *params)

    @torch.no_grad()
    def _forward_with_record(self, events, record: bool):
        """The reference's analysis outputs (embedding.py:180, 198-199, 221-224; the Fig. 4 study, readme.md:162):
        ``record`` -> (frames, t_last history ``[steps, B, 2, H, W]`` int64), ``v_record`` -> (frames, the
        sub-threshold potentials of every step, concatenated).  Both are derived from the per-step potentials the
        kernel saves for training (``v_seq``); the bookkeeping is replayed with a few tensor ops (analysis path)."""
        out, v_seq, _, _, _ = _sampler_forward(events, self, self._params(), True)
        Tm = v_seq.shape[0]
        spikes = (v_seq - float(self.thresh)) > 0                      # Rectangle.forward, activation.py:21-23
        seg = torch.zeros_like(v_seq[0], dtype=torch.int64)
        t_last = torch.full_like(seg, -1)
        t_rec, v_rec = [], []
        for t in range(Tm):
            v_rec.append(v_seq[t][~spikes[t]])
            valid = spikes[t] & (seg < self.Ts)
            seg = seg + valid.long()
            t_last = torch.where(valid, torch.full_like(t_last, t), t_last)
            t_rec.append(t_last.clone())
            if int(seg.min()) >= self.Ts:                              # the reference's early break (:200-201)
                break
        if record:
            return out, torch.stack(t_rec, dim=0)
        return out, torch.cat(v_rec)

    def _hist_dtype(self, B, H, W, strategy, hist_dtype):
        """The histogram format between binning and sampling: ``"compact"`` = one byte per bin + the exact list of
        saturated bins (a quarter of the bytes written and read; the tensor-core kernel and the tiles binning
        kernel), ``"dense"`` = fp32 counts, ``"auto"`` = compact whenever those two kernels take the call and the batch
        gives the tiles kernel a work item per SM (a lone window bins faster through the global-reduction kernel)."""
        if hist_dtype not in ("auto", "compact", "dense"):
            raise ValueError("hist_dtype must be 'auto', 'compact' or 'dense'")
        ok = (self.depth == 2 and self.kernel_size == 5 and W % 4 == 0 and self.Ts <= 15 and self.nb_steps <= 14 and
              self.algo in ("auto", "tensor") and strategy in ("auto", "tiles") and compact_fits(H, W) and
              not (torch.is_grad_enabled() and any(q.requires_grad for q in self.parameters())))
        if hist_dtype == "compact" and not ok:
            raise ValueError("the compact histogram needs the tensor-core sampler (depth 2, k 5, W % 4 == 0, inference) "
                             "and a frame that fits the tiles binning kernel")
        if hist_dtype == "auto" and ok:
            ok = B * self.nb_steps * 2 * compact_fits(H, W) >= 148
        return torch.uint8 if (ok and hist_dtype != "dense") else torch.float32

    def forward_events(self, x, y, t, p, offsets, H: int, W: int, strategy: str = "auto", hist_dtype: str = "auto"):
        """Raw time-sorted event windows -> adaptive frames ``[Ts, B, 2, H, W]``.

        Binning (gen1.py:313-360) and sampling (embedding.py:141-226) back to back on the GPU; the histogram between
        them is the compact byte form (:class:`eas_snn_b200.binning.CompactHist`) where the kernels support it, fp32
        counts otherwise -- the same frames either way.
        """
        hist = bin_events(x, y, t, p, offsets, H, W, self.nb_steps, strategy=strategy,
                          dtype=self._hist_dtype(offsets.numel() - 1, H, W, strategy, hist_dtype))
        return self.forward(hist)

    def forward_dat(self, records, ranges, H: int, W: int, strategy: str = "auto", hist_dtype: str = "auto"):
        """Raw PSEE ``.dat`` records + one record range per window (:func:`eas_snn_b200.psee.dat_windows`)
        -> adaptive frames ``[Ts, B, 2, H, W]``: decode, binning and sampling without leaving the GPU."""
        from .psee import bin_dat
        hist = bin_dat(records, ranges, H, W, self.nb_steps, strategy=strategy,
                       dtype=self._hist_dtype(ranges.shape[0], H, W, strategy, hist_dtype))
        return self.forward(hist)


class SpikeCountEmbedding(nn.Module):
    """``SpikeCountEmbedding`` (yolox/models/embedding.py:9-24; ``embedding: count`` in event_yolox_base.py:160): the
    micro-bin histograms of each window summed over the Tm micro-bins -- the plain event-count frame the adaptive
    sampler is compared against.  ``forward(events)`` takes what the reference takes: 5-D ``[B, Tm, 2, H, W]`` ->
    ``[B, 2, H, W]``, 6-D ``[B, Tl, Tm, 2, H, W]`` -> ``[B*Tl, 2, H, W]``, and < 5-D (a single frame) ->
    ``nb_steps`` copies summed = ``nb_steps * events``.  No parameters."""

    def __init__(self, nb_steps):
        super().__init__()
        self.nb_steps = nb_steps

    def forward(self, events: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda(events)
        if events.dim() < 5:                       # embedding.py:15-16: broadcast over nb_steps, then sum
            return events.float() * float(self.nb_steps)
        ev = events.flatten(end_dim=-5) if events.dim() > 5 else events
        if ev.dtype not in (torch.float32, torch.int32):
            ev = ev.float()
        ev = ev.contiguous()
        n, Tm = ev.shape[0], ev.shape[1]
        plane = ev[0, 0].numel()
        if plane % 4 != 0:
            return ev.float().sum(dim=1)           # ragged planes: not worth a kernel
        out = torch.empty(ev.shape[:1] + ev.shape[2:], dtype=torch.float32, device=ev.device)
        in_dtype = _lib.EAS_I32 if ev.dtype == torch.int32 else _lib.EAS_F32
        with torch.cuda.device(ev.device):
            rc = _lib.lib().eas_hist_time_sum(_lib.ptr(ev), in_dtype, n, Tm, plane, _lib.ptr(out), _lib.stream_ptr())
        _lib.check(rc, "eas_hist_time_sum")
        return out

    def forward_events(self, x, y, t, p, offsets, H: int, W: int):
        """Raw windows -> count frames ``[B, 2, H, W]``: binning (gen1.py:313-360) + the sum over micro-bins."""
        from .binning import CompactHist, bin_events, compact_fits
        return self.forward(bin_events(x, y, t, p, offsets, H, W, self.nb_steps))


class _SeqReadoutEmbedding(nn.Module):
    """Shared host side of the two ablation embeddings: both are the sampler's recurrence without the
    spike-triggered aggregation, read out as the sum over the steps of the pre-reset potential (``readout='sum'``)
    or as the final membrane potential (``'last'``).  They run on ``eas_sampler_fwd`` and read the per-step potentials
    it saves (``v_seq``).  Inference only (the surrogate backward of these read-outs is not built)."""

    Ts = 1
    spike_attach = False
    write_zero = True
    abs = False
    algo = "auto"

    def _prep(self, events):
        if events.dim() < 5:   # parameter-registering input: nb_steps copies of one frame (embedding.py:44-45, :287-288)
            events = events.unsqueeze(0).expand((self.nb_steps,) + tuple(events.shape)).transpose(0, 1)
        elif events.dim() > 5:
            events = events.flatten(end_dim=-5)
        _lib.require_cuda(events)
        if events.shape[2] != 2:
            raise ValueError("expected [B, Tm, 2, H, W] micro-bin tensor")
        if events.dtype not in (torch.float32, torch.int32):
            events = events.float()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and self.training:
            raise NotImplementedError("%s: inference only on the sm_100a path" % type(self).__name__)
        return events.contiguous()

    def _readout(self, v_seq, relu=False):
        if self.readout == "sum":
            agg = v_seq[0].clone()
            for t in range(1, v_seq.shape[0]):        # same order as the reference's running sum
                agg += v_seq[t]
        elif self.readout == "last":
            v = v_seq[-1]
            s = ((v - float(self.thresh)) > 0).float()
            agg = v - float(self.thresh) * s if self.vreset is None else v * (1 - s) + float(self.vreset) * s
        else:
            raise NotImplementedError(self.readout)
        return torch.relu(agg) if relu else agg


class SpikingEmbedding(_SeqReadoutEmbedding):
    """``SpikingEmbedding`` (``embedding: rsnn``; yolox/models/embedding.py:229-316): the gated recurrent spiking layer
    of the adaptive sampler (same ``update``, :276-283; same convolutions, the input stack wrapped in ``tdLayer`` ->
    keys ``input_conv.layer.{0,2}.*``) whose output is sum_t v_t or the last membrane potential, ``[B*Tl, 2, H, W]``."""

    def __init__(self, kernel_size, in_channel=2, out_channel=2, readout="sum", relu=False, depth=1, **kwargs_spikes):
        super().__init__()
        if in_channel != 2 or out_channel != 2 or int(depth) not in (1, 2) or int(kernel_size) not in (3, 5, 7):
            raise NotImplementedError("sampler kernels: 2 polarity channels, depth in {1,2}, kernel_size in {3,5,7}")
        fn = kwargs_spikes.get("spike_fn", None)
        if fn is not None and getattr(fn, "__name__", type(fn).__name__) != "Rectangle":
            raise NotImplementedError("the sampler kernel implements the Rectangle spike function")
        self.kernel_size, self.depth, self.readout, self.relu = int(kernel_size), int(depth), readout, relu
        self.kwargs_spikes = kwargs_spikes
        self.nb_steps = kwargs_spikes["nb_steps"] if "Tm" not in kwargs_spikes else kwargs_spikes["Tm"]
        self.thresh = kwargs_spikes["thresh"]
        self.vreset = copy.deepcopy(kwargs_spikes["vreset"])
        self.input_conv = _TdLayer(AdaptiveRSNNEmbedding.build_conv(in_channel, out_channel * 2, self.kernel_size, self.depth))
        self.gate_conv = AdaptiveRSNNEmbedding.build_conv(out_channel, out_channel * 2, self.kernel_size, self.depth)
        for m in self.input_conv.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.orthogonal_(m.weight, gain=nn.init.calculate_gain("relu"))
        for m in self.gate_conv.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, nonlinearity="sigmoid")

    def forward(self, events):
        events = self._prep(events)
        params = []
        for convs in (self.input_conv.layer, self.gate_conv):
            for m in convs:
                if isinstance(m, nn.Conv2d):
                    params += [m.weight, m.bias]
        with torch.no_grad():
            _, v_seq, _, _, _ = _sampler_forward(events, self, params, True)
            return self._readout(v_seq, relu=self.relu)


class LIFEmbedding(_SeqReadoutEmbedding):
    """``LIFEmbedding`` (``embedding: snn``; yolox/models/embedding.py:28-76): a feed-forward conv stack (``tdLayer``,
    keys ``embedding_conv.layer.{0,2}.*``) into one ``LIFCell`` (cell.py:37-65: ``v = sigmoid(decay) * v + psp``,
    Rectangle spike, soft / hard reset), read out as sum_t v_t or the last potential.  This is the sampler's recurrence
    with a constant gate and no recurrent input, so it runs on the sampler kernels with synthesised weights: gate
    channels = zero weights + bias ``decay`` (their sigmoid is the cell's leak), recurrent stack = zeros, current
    channels = the embedding convolution."""

    def __init__(self, kernel_size, in_channel=2, out_channel=2, readout="sum", depth=1, **kwargs_spikes):
        super().__init__()
        if in_channel != 2 or out_channel != 2 or int(depth) not in (1, 2) or int(kernel_size) not in (3, 5, 7):
            raise NotImplementedError("sampler kernels: 2 polarity channels, depth in {1,2}, kernel_size in {3,5,7}")
        fn = kwargs_spikes.get("spike_fn", None)
        if fn is not None and getattr(fn, "__name__", type(fn).__name__) != "Rectangle":
            raise NotImplementedError("the sampler kernel implements the Rectangle spike function")
        self.kernel_size, self.depth, self.readout = int(kernel_size), int(depth), readout
        self.kwargs_spikes = kwargs_spikes
        self.nb_steps = kwargs_spikes["nb_steps"] if "Tm" not in kwargs_spikes else kwargs_spikes["Tm"]
        self.embedding_conv = _TdLayer(AdaptiveRSNNEmbedding.build_conv(in_channel, out_channel, self.kernel_size, self.depth))
        self.cell = _LIFCellParams(decay=kwargs_spikes.get("decay", None), thresh=kwargs_spikes.get("thresh", None),
                                   vreset=kwargs_spikes.get("vreset", None))
        for m in self.embedding_conv.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.orthogonal_(m.weight, gain=nn.init.calculate_gain("relu"))

    @property
    def thresh(self):
        return self.cell.thresh

    @property
    def vreset(self):
        return self.cell.vreset

    def forward(self, events):
        events = self._prep(events)
        convs = [m for m in self.embedding_conv.layer if isinstance(m, nn.Conv2d)]
        k, dev = self.kernel_size, events.device
        decay = self.cell.decay.detach().float().reshape(()).to(dev)
        with torch.no_grad():
            def widen(conv, cin_pad, last):
                """Conv(c -> 2) as the sampler's Conv(cin_pad -> 4): [gate pair | current pair] output channels (last layer)
                or [hidden pair | unused pair] (first of two layers)."""
                w = torch.zeros((4, cin_pad, k, k), device=dev)
                b = torch.zeros((4,), device=dev)
                o = 2 if last else 0
                w[o:o + 2, :conv.weight.shape[1]] = conv.weight.float()
                b[o:o + 2] = conv.bias.float()
                if last:
                    b[0:2] = decay                      # gate pre-activation = decay: sigmoid(decay) is the leak
                return w, b
            if self.depth == 1:
                iw0, ib0 = widen(convs[0], 2, True)
                params = [iw0, ib0, torch.zeros_like(iw0), torch.zeros_like(ib0)]
            else:
                iw0, ib0 = widen(convs[0], 2, False)
                iw1, ib1 = widen(convs[1], 4, True)
                z0, z1 = torch.zeros_like(iw0), torch.zeros_like(iw1)
                params = [iw0, ib0, iw1, ib1, z0, torch.zeros_like(ib0), z1, torch.zeros_like(ib1)]
            _, v_seq, _, _, _ = _sampler_forward(events, self, params, True)
            return self._readout(v_seq)


class _TdLayer(nn.Module):
    """``tdLayer`` (yolox/models/layer.py:122-132): only its child name (``layer``) matters here."""

    def __init__(self, layer):
        super().__init__()
        self.layer = layer


class _LIFCellParams(nn.Module):
    """The parameters of ``LIFCell`` (cell.py:24-36, 67-79): ``decay`` (a logit; an ``nn.Parameter`` when the caller
    passes one, as ``EventExp.get_kwargs_spikes`` does), ``thresh``, ``vreset``."""

    def __init__(self, decay=None, thresh=None, vreset=None):
        super().__init__()
        self.decay = copy.deepcopy(decay) if decay is not None else nn.Parameter(torch.tensor(0.0))
        if not isinstance(self.decay, torch.Tensor):
            self.decay = torch.tensor(float(self.decay))
        self.thresh = 0.5 if thresh is None else copy.deepcopy(thresh)
        self.vreset = copy.deepcopy(vreset)
