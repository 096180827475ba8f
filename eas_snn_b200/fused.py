"""conv -> BatchNorm -> multi-step PLIF fused on the tensor cores, and the spiking CSPDarknet built
from it.

Drop-in seam: the reference turns every ``BaseConv`` into
``SeqToANNContainer(Conv2d) -> layer.BatchNorm2d('m') -> ParametricLIFNode``
(``yolox/utils/utils_snn.py:16-58``, ``yolox/models/network_blocks.py:31-56``).  :class:`FusedConvBNPLIF`
keeps those three children and their state-dict keys (``conv.0.weight``, ``bn.*``, ``act.w``) but in
eval mode runs ONE kernel (``eas_conv_bn_plif_fwd``): implicit-GEMM conv on tcgen05 with the BN
folded into fp16-split weights (``yolox/utils/model_utils.py:61-75``) and the LIF recurrence over T in
the epilogue, so neither the conv output nor the membrane potential reaches HBM.

Activations between fused layers are channels-last fp16 ``[T, B, H, W, C]`` (spikes and SEW sums
are small integers, exact in fp16).  :class:`SpikingCSPDarknet` is the reference topology
(``yolox/models/darknet.py:97-180``) executed on that layout, with concatenations realised by
writing conv outputs straight into channel slices of the concat buffer.
"""
from __future__ import annotations

import copy
import os
import ctypes as C

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .neuron import ATan, ParametricLIFNode

OUT_SPIKES, OUT_PREACT, OUT_SILU2 = 0, 1, 2
ACT_DTYPE = torch.float16      # activations, weight planes


# ------------------------------------------------------------------------------------------------
# functional layer
# ------------------------------------------------------------------------------------------------
def split_f16(x: torch.Tensor, n: int = 2) -> torch.Tensor:
    """fp32 tensor -> ``[n, ...]`` fp16 planes whose (fp32) sum reproduces x to ~2^-(11n) (|x| < 65504)."""
    planes, r = [], x.float().clamp(-65504.0, 65504.0)
    for _ in range(n):
        p = r.to(torch.float16)
        planes.append(p)
        r = r - p.float()
    return torch.stack(planes)


def fold_bn(conv_w: torch.Tensor, bn_w, bn_b, mean, var, eps: float):
    """``fuse_conv_and_bn`` (yolox/utils/model_utils.py:61-75) in fp32: (w', shift)."""
    scale = bn_w.float() / torch.sqrt(var.float() + eps)
    return conv_w.float() * scale.view(-1, 1, 1, 1), bn_b.float() - mean.float() * scale


def pack_weight(w_folded: torch.Tensor, n_wsplit: int = 2):
    """``[Cout, Cin, kh, kw]`` fp32 -> (``[n_wsplit, Cout, kh, kw, Cin]`` fp16 planes, K-major for TMA;
    ``[Cout]`` fp32 unscale).  Every output channel is scaled by a power of two (exact) that puts its
    largest weight in [2^13, 2^14), so the hi and lo planes are normal fp16 numbers; the kernel multiplies
    the accumulator by ``unscale`` = the inverse power of two before adding the bias."""
    w = w_folded.float()
    amax = w.abs().amax(dim=(1, 2, 3)).clamp_min(2.0 ** -100)
    e = (13 - torch.floor(torch.log2(amax))).clamp(-24, 40)
    scale = torch.exp2(e)
    planes = split_f16((w * scale.view(-1, 1, 1, 1)).permute(0, 2, 3, 1).contiguous(), n_wsplit).contiguous()
    return planes, torch.exp2(-e).contiguous()


def _pixel_ld(t: torch.Tensor, C: int) -> int:
    """Pixel stride (elements) of a channels-last ``[..., B, H, W, C]`` tensor that may be a channel slice of a
    wider buffer: read off the innermost dimension that has more than one entry."""
    n = 1
    for d in range(t.dim() - 2, -1, -1):
        if t.shape[d] > 1:
            return t.stride(d) // n
        n *= t.shape[d]
    return C


def conv_bn_plif(x: torch.Tensor, w_planes: torch.Tensor, bias: torch.Tensor, plif_w: torch.Tensor | None,
                 T: int, ksize: int, stride: int, n_xsplit: int = 1, out: torch.Tensor | None = None,
                 out_mode: int = OUT_SPIKES, v_threshold: float = 1.0, v_reset: float | None = None,
                 decay_input: bool = False, residual: torch.Tensor | None = None,
                 w_unscale: torch.Tensor | None = None) -> torch.Tensor:
    """One fused layer.

    x        : channels-last fp16 ``[Tx, B, H, W, Cin]`` (``[n_xsplit, Tx, B, H, W, Cin]`` for split
               real-valued input); may be a channel slice of a wider buffer (pixel stride ``x_ld``).
    w_planes : ``[n_wsplit, Cout, k, k, Cin]`` fp16 and ``w_unscale`` ``[Cout]`` fp32 (:func:`pack_weight`),
               bias ``[Cout]`` fp32.
    out      : optional destination view ``[T, B, Ho, Wo, Cout]`` (channel slice of a concat buffer).
    residual : optional fp16 ``[T, B, Ho, Wo, Cout]`` added to the spikes in the epilogue (SEW shortcut).
    Returns spikes ``[T, B, Ho, Wo, Cout]`` fp16 (or the fp32 pre-activation / 2 SiLU planes).
    """
    _lib.require_cuda(x, w_planes, bias, w_unscale)
    if x.dtype != ACT_DTYPE or w_planes.dtype != ACT_DTYPE:
        raise TypeError("conv_bn_plif expects fp16 activations and weight planes")
    if w_unscale is not None and (w_unscale.dtype != torch.float32 or w_unscale.numel() != w_planes.shape[1]):
        raise TypeError("w_unscale must be fp32 [Cout]")
    xs = x if n_xsplit > 1 or x.dim() == 6 else x.unsqueeze(0)
    if xs.dim() != 6 or xs.shape[0] != n_xsplit:
        raise ValueError("x must be [Tx,B,H,W,C] or [n_xsplit,Tx,B,H,W,C]")
    _, Tx, B, H, W, Cin = xs.shape
    x_ld = _pixel_ld(xs, Cin)
    want = (Tx * B * H * W * x_ld, B * H * W * x_ld, H * W * x_ld, W * x_ld, x_ld, 1)
    if x_ld < Cin or any(d > 1 and s_ != w_ for s_, w_, d in zip(xs.stride(), want, xs.shape)):
        raise ValueError("x must be a dense channels-last tensor or a channel slice of one "
                         "(shape %s, strides %s)" % (tuple(xs.shape), tuple(xs.stride())))
    n_wsplit, Cout, kh, kw, Cin_w = w_planes.shape
    if Cin_w != Cin or kh != ksize or kw != ksize:
        raise ValueError("weight planes do not match the input")
    pad = (ksize - 1) // 2
    Ho = (H + 2 * pad - ksize) // stride + 1
    Wo = (W + 2 * pad - ksize) // stride + 1
    dev = x.device
    To = T if out_mode == OUT_SPIKES else Tx
    if out is None:
        if out_mode == OUT_SPIKES:
            out = torch.empty((To, B, Ho, Wo, Cout), dtype=ACT_DTYPE, device=dev)
        elif out_mode == OUT_PREACT:
            out = torch.empty((To, B, Ho, Wo, Cout), dtype=torch.float32, device=dev)
        else:
            out = torch.empty((2, To, B, Ho, Wo, Cout), dtype=ACT_DTYPE, device=dev)
    if out_mode == OUT_SILU2 and (out.dim() != 6 or out.shape[0] != 2):
        raise ValueError("OUT_SILU2 writes two planes: out must be [2, Tx, B, Ho, Wo, Cout]")
    oshape = tuple(out.shape[-5:])
    if oshape != (To, B, Ho, Wo, Cout) or (Cout > 1 and out.stride(-1) != 1):
        raise ValueError("out has the wrong shape %s, expected %s" % (oshape, (To, B, Ho, Wo, Cout)))
    out_ld = _pixel_ld(out, Cout)
    owant = (B * Ho * Wo * out_ld, Ho * Wo * out_ld, Wo * out_ld, out_ld, 1)
    if out_ld < Cout or any(d > 1 and s_ != w_ for s_, w_, d in zip(out.stride()[-5:], owant, oshape)):
        raise ValueError("out must be a dense channels-last tensor or a channel slice of one")
    if out_mode == OUT_SILU2 and out.stride(0) != To * B * Ho * Wo * out_ld:
        raise ValueError("the two output planes must be one whole buffer apart")
    res_ld = 0
    if residual is not None:
        if residual.dtype != ACT_DTYPE or tuple(residual.shape) != (T, B, Ho, Wo, Cout):
            raise ValueError("residual must be fp16 [T, B, Ho, Wo, Cout]")
        res_ld = _pixel_ld(residual, Cout)
        rwant = (B * Ho * Wo * res_ld, Ho * Wo * res_ld, Wo * res_ld, res_ld, 1)
        if any(d > 1 and s_ != w_ for s_, w_, d in zip(residual.stride(), rwant, residual.shape)):
            raise ValueError("residual must be channels-last (or a channel slice)")
    cfg = _lib.ConvCfg(T=T, Tx=Tx, B=B, H=H, W=W, Cin=Cin, Cout=Cout, ksize=ksize, stride=stride,
                       n_wsplit=n_wsplit, n_xsplit=n_xsplit, v_threshold=float(v_threshold),
                       hard_reset=0 if v_reset is None else 1, v_reset=0.0 if v_reset is None else float(v_reset),
                       decay_input=int(bool(decay_input)), out_mode=out_mode, x_ld=x_ld, out_ld=out_ld,
                       res_ld=res_ld, residual=None if residual is None else residual.data_ptr(),
                       w_unscale=None if w_unscale is None else w_unscale.contiguous().data_ptr())
    L = _lib.lib()
    with torch.cuda.device(dev):
        rc = L.eas_conv_bn_plif_fwd(C.byref(cfg), _lib.ptr(xs), _lib.ptr(w_planes), _lib.ptr(bias),
                                    _lib.ptr(plif_w), _lib.ptr(out), None, 0, _lib.stream_ptr())
    _lib.check(rc, "eas_conv_bn_plif_fwd")
    return out


def to_channels_last_f16(x_seq: torch.Tensor) -> torch.Tensor:
    """``[T, B, C, H, W]`` (any strides / float dtype) -> ``[T, B, H, W, C]`` fp16 contiguous.
    Free when x_seq is already a permuted view of such a buffer (what the fused layers return)."""
    y = x_seq.permute(0, 1, 3, 4, 2)
    if y.dtype != ACT_DTYPE:
        y = y.to(ACT_DTYPE)
    return y.contiguous()


# ------------------------------------------------------------------------------------------------
# module with the reference's children / state-dict keys
# ------------------------------------------------------------------------------------------------
class SeqToANNContainer(nn.Sequential):
    """``spikingjelly.activation_based.layer.SeqToANNContainer``: flatten [T, B] -> apply -> restore."""

    def forward(self, x_seq):
        y = super().forward(x_seq.flatten(0, 1))
        return y.view([x_seq.shape[0], x_seq.shape[1]] + list(y.shape[1:]))


class MultiStepBatchNorm2d(nn.BatchNorm2d):
    """``layer.BatchNorm2d(..., step_mode='m')``: statistics over the flattened T*B batch.

    ``count_in_forward = False`` (set by a parent that counts for all of its layers at once, ``bump_bn_counters``):
    the training forward leaves ``num_batches_tracked`` alone -- with a fixed ``momentum`` the counter does not enter the
    result, and 35 one-element ``add_`` kernels per step are 1.6 % of a SYOLOX-S training step."""
    step_mode = "m"
    count_in_forward = True

    def forward(self, x):
        if x.dim() != 5:
            raise ValueError("expected [T, N, C, H, W]")
        x4 = x.flatten(0, 1)
        if self.training and self.track_running_stats and self.momentum is not None and not self.count_in_forward:
            y = F.batch_norm(x4, self.running_mean, self.running_var, self.weight, self.bias, True, self.momentum,
                             self.eps)
        else:
            y = super().forward(x4)
        return y.view([x.shape[0], x.shape[1]] + list(y.shape[1:]))


def bump_bn_counters(model: nn.Module):
    """``num_batches_tracked += 1`` for every ``MultiStepBatchNorm2d`` of ``model`` in ONE multi-tensor kernel, and from
    now on those layers do not count in their own forward (call once per training forward)."""
    layers = getattr(model, "_bn_counted", None)
    if layers is None:
        layers = [m for m in model.modules()
                  if isinstance(m, MultiStepBatchNorm2d) and m.track_running_stats and m.momentum is not None]
        for m in layers:
            m.count_in_forward = False
        object.__setattr__(model, "_bn_counted", layers)     # (a plain attribute: not a registered sub-module list)
    if layers:
        torch._foreach_add_([m.num_batches_tracked for m in layers], 1)   # (buffers looked up now: .to() replaces them)


class FusedConvBNPLIF(nn.Module):
    """``BaseConv`` after ``convert_to_spiking`` (network_blocks.py:31-56, utils_snn.py:25-53)."""

    def __init__(self, in_channels, out_channels, ksize, stride, spike_fn=None, n_wsplit: int = 2,
                 eps: float = 1e-3, momentum: float = 0.03):
        super().__init__()
        self.conv = SeqToANNContainer(nn.Conv2d(in_channels, out_channels, ksize, stride, (ksize - 1) // 2,
                                                bias=False))
        self.bn = MultiStepBatchNorm2d(out_channels, eps=eps, momentum=momentum)
        self.act = ParametricLIFNode(init_tau=2.0, decay_input=False, v_threshold=1.0, v_reset=None,
                                     surrogate_function=copy.deepcopy(spike_fn) if spike_fn is not None else ATan(2.0),
                                     detach_reset=False, step_mode="m", backend="torch")
        self.ksize, self.stride, self.n_wsplit = ksize, stride, n_wsplit
        self.assume_integer_input = False   # set True when the input is known to be spikes / SEW sums
        self._cache = None

    @classmethod
    def from_modules(cls, conv: nn.Conv2d, bn: nn.BatchNorm2d, act=None, spike_fn=None, n_wsplit: int = 2):
        if conv.groups != 1 or conv.bias is not None or conv.kernel_size[0] != conv.kernel_size[1]:
            raise NotImplementedError("fused layer covers the reference's BaseConv: square, groups=1, no bias")
        m = cls(conv.in_channels, conv.out_channels, conv.kernel_size[0], conv.stride[0], spike_fn=spike_fn,
                n_wsplit=n_wsplit, eps=bn.eps, momentum=bn.momentum)
        m.conv[0].load_state_dict(conv.state_dict())
        m.bn.load_state_dict(bn.state_dict())
        if isinstance(act, ParametricLIFNode):
            m.act = act
        return m.to(conv.weight.device)

    # -- folded / split weights, rebuilt when any source tensor changes --------------------------
    def _sources(self):
        return (self.conv[0].weight, self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var)

    def packed(self):
        # (bn.eps is part of the key: init_yolo rewrites it after construction, event_yolox_base.py:179-183)
        key = tuple((t.data_ptr(), t._version) for t in self._sources()) + (self.n_wsplit, float(self.bn.eps))
        if self._cache is None or self._cache[0] != key:
            with torch.no_grad():
                w, shift = fold_bn(self.conv[0].weight, self.bn.weight, self.bn.bias, self.bn.running_mean,
                                   self.bn.running_var, self.bn.eps)
                wp, unscale = pack_weight(w, self.n_wsplit)
                self._cache = (key, wp, shift.contiguous(), unscale)
        return self._cache[1], self._cache[2], self._cache[3]

    # -- channels-last fast path (used by SpikingCSPDarknet) -------------------------------------
    def run(self, x_cl: torch.Tensor, T: int, out: torch.Tensor | None = None, n_xsplit: int = 1,
            residual: torch.Tensor | None = None) -> torch.Tensor:
        wp, shift, unscale = self.packed()
        a = self.act
        return conv_bn_plif(x_cl, wp, shift, a.w.detach().float(), T, self.ksize, self.stride, n_xsplit=n_xsplit,
                            out=out, out_mode=OUT_SPIKES, v_threshold=a.v_threshold, v_reset=a.v_reset,
                            decay_input=a.decay_input, residual=residual, w_unscale=unscale)

    # -- reference-shaped forward: [T, B, C, H, W] in, [T, B, C', H', W'] out -----------------------
    def forward(self, x_seq: torch.Tensor) -> torch.Tensor:
        if self.training:
            # training keeps batch statistics and autograd: conv and BN through PyTorch, the neuron
            # (forward and surrogate backward) through the fused PLIF kernels
            return self.act(self.bn(self.conv(x_seq)))
        T = x_seq.shape[0]
        x_cl = x_seq.permute(0, 1, 3, 4, 2)
        if x_seq.dtype in (torch.float16, torch.bfloat16) or self.assume_integer_input:
            y = self.run(x_cl.to(ACT_DTYPE).contiguous(), T)     # spikes / SEW sums: exact in fp16
        else:  # real-valued input (e.g. the SiLU stem output): split it so the product stays fp32-accurate
            y = self.run(split_f16(x_cl.contiguous(), 2), T, n_xsplit=2)
        self.act.reset()                                   # stateless: the reference resets per batch
        return y.permute(0, 1, 4, 2, 3).to(x_seq.dtype)    # logical [T, B, C, H, W], channels-last memory


# ------------------------------------------------------------------------------------------------
# the spiking CSPDarknet on channels-last buffers
# ------------------------------------------------------------------------------------------------
class AnnBaseConv(nn.Module):
    """The reference's ANN ``BaseConv``: conv -> BN -> SiLU (network_blocks.py:31-56); keys ``conv.weight``, ``bn.*``.
    Used for the stem's inner conv (``Focus`` stays ANN, utils_snn.py:23-24) and for every layer of the ANN pyramid /
    head (``eas_snn_b200.detector``).  ``run`` works on two-plane activations ``[2, 1, B, H, W, C]``."""

    def __init__(self, cin, cout, ksize, stride=1):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, ksize, stride, (ksize - 1) // 2, bias=False)
        self.bn = nn.BatchNorm2d(cout, eps=1e-3, momentum=0.03)       # init_yolo, event_yolox_base.py:179-183
        self.act = nn.SiLU()
        self.ksize, self.stride = ksize, stride
        self.fp16_inputs = False      # True: read only the hi plane of the input (SpikingYOLOX.set_ann_precision)
        self._cache = None

    def packed(self):
        src = (self.conv.weight, self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var)
        key = tuple((t.data_ptr(), t._version) for t in src) + (float(self.bn.eps),)
        if self._cache is None or self._cache[0] != key:
            with torch.no_grad():
                w, shift = fold_bn(self.conv.weight, self.bn.weight, self.bn.bias, self.bn.running_mean,
                                   self.bn.running_var, self.bn.eps)
                wp, unscale = pack_weight(w, 2)
                self._cache = (key, wp, shift.contiguous(), unscale)
        return self._cache[1], self._cache[2], self._cache[3]

    def run(self, xp: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        wp, shift, unscale = self.packed()
        if self.fp16_inputs:          # fp16 activations x fp32-equivalent weights: two product terms instead of three
            return conv_bn_plif(xp[:1], wp, shift, None, 1, self.ksize, self.stride, n_xsplit=1, out=out,
                                out_mode=OUT_SILU2, w_unscale=unscale)
        return conv_bn_plif(xp, wp, shift, None, 1, self.ksize, self.stride, n_xsplit=2, out=out, out_mode=OUT_SILU2,
                            w_unscale=unscale)


class _Focus(nn.Module):
    def __init__(self, cin, cout, k):
        super().__init__()
        self.conv = AnnBaseConv(cin * 4, cout, k, 1)

    def forward(self, x):
        """PyTorch path (training: batch statistics + autograd), ``[N, C, H, W]`` (network_blocks.py:199-213)."""
        a, b = x[..., ::2, ::2], x[..., 1::2, ::2]
        c, d = x[..., ::2, 1::2], x[..., 1::2, 1::2]
        m = self.conv
        return m.act(m.bn(m.conv(torch.cat((a, b, c, d), dim=1))))

    def packed_im2col(self):
        """The stem conv as a 1x1 conv over im2col rows ``[tap][focus channel]`` (72 -> 80 zero padded)."""
        m = self.conv
        src = (m.conv.weight, m.bn.weight, m.bn.bias, m.bn.running_mean, m.bn.running_var)
        key = tuple((t.data_ptr(), t._version) for t in src) + (float(m.bn.eps),)
        if getattr(self, "_cache_i2c", None) is None or self._cache_i2c[0] != key:
            with torch.no_grad():
                w, shift = fold_bn(m.conv.weight, m.bn.weight, m.bn.bias, m.bn.running_mean, m.bn.running_var, m.bn.eps)
                w80 = torch.zeros((w.shape[0], 80, 1, 1), dtype=torch.float32, device=w.device)
                w80[:, :72, 0, 0] = w.permute(0, 2, 3, 1).reshape(w.shape[0], 72)       # (ky, kx, cin) order
                wp, unscale = pack_weight(w80, 2)
                self._cache_i2c = (key, wp, shift.contiguous(), unscale)
        return self._cache_i2c[1:]

    def run(self, frames: torch.Tensor) -> torch.Tensor:
        """frames ``[Tx, B, C, H, W]`` fp32 -> SiLU(BN(conv(space_to_depth))) as 2 fp16 planes
        ``[2, Tx, B, H/2, W/2, cout]`` (network_blocks.py:191-213)."""
        Tx, B, Cf, H, W = frames.shape
        if Cf == 2 and self.conv.conv.kernel_size == (3, 3) and H % 2 == 0 and W % 2 == 0:
            # space-to-depth + im2col + fp16 hi/lo split in one pass, then a K = 80 1x1 conv on the tensor cores
            fr = frames.float().contiguous()
            cols = torch.empty((2, Tx, B, H // 2, W // 2, 80), dtype=ACT_DTYPE, device=frames.device)
            with torch.cuda.device(frames.device):
                rc = _lib.lib().eas_focus_im2col(_lib.ptr(fr), Tx * B, H, W, _lib.ptr(cols), cols.stride(0),
                                                 _lib.stream_ptr())
            _lib.check(rc, "eas_focus_im2col")
            wp, shift, unscale = self.packed_im2col()
            return conv_bn_plif(cols, wp, shift, None, Tx, 1, 1, n_xsplit=2, out_mode=OUT_SILU2, w_unscale=unscale)
        a, b = frames[..., ::2, ::2], frames[..., 1::2, ::2]
        c, d = frames[..., ::2, 1::2], frames[..., 1::2, 1::2]
        x = torch.cat((a, b, c, d), dim=2).permute(0, 1, 3, 4, 2).contiguous()      # channels-last fp32
        wp, shift, unscale = self.conv.packed()
        return conv_bn_plif(split_f16(x, 2), wp, shift, None, x.shape[0], 3, 1, n_xsplit=2, out_mode=OUT_SILU2,
                            w_unscale=unscale)


class _Bottleneck(nn.Module):
    def __init__(self, cin, cout, shortcut, expansion, spike_fn):
        super().__init__()
        hid = int(cout * expansion)
        self.conv1 = FusedConvBNPLIF(cin, hid, 1, 1, spike_fn)
        self.conv2 = FusedConvBNPLIF(hid, cout, 3, 1, spike_fn)
        self.use_add = shortcut and cin == cout

    def run(self, x, T, out=None):
        # SEW add y + x (network_blocks.py:99-103) happens in the conv epilogue: values {0,1,2,..}
        return self.conv2.run(self.conv1.run(x, T), T, out=out, residual=x if self.use_add else None)

    def forward(self, x):          # reference-shaped [T, B, C, H, W] (training path of the layers)
        y = self.conv2(self.conv1(x))
        return y + x if self.use_add else y


class _SPP(nn.Module):
    def __init__(self, cin, cout, spike_fn, ks=(5, 9, 13)):
        super().__init__()
        hid = cin // 2
        self.ks = ks
        self.conv1 = FusedConvBNPLIF(cin, hid, 1, 1, spike_fn)
        self.m = nn.ModuleList([SeqToANNContainer(nn.MaxPool2d(k, 1, k // 2)) for k in ks])
        self.conv2 = FusedConvBNPLIF(hid * (len(ks) + 1), cout, 1, 1, spike_fn)

    def run(self, x, T):
        Tn, B, H, W, _ = x.shape
        hid = self.conv1.conv[0].out_channels
        cat = torch.empty((T, B, H, W, hid * (len(self.ks) + 1)), dtype=ACT_DTYPE, device=x.device)
        y = self.conv1.run(x, T, out=cat[..., :hid])
        if len(self.ks) == 3 and hid % 8 == 0 and H * W * 256 <= 200 * 1024:
            with torch.cuda.device(x.device):          # the three pools in one pass, into the concat slices
                rc = _lib.lib().eas_spp_pool_fwd(_lib.ptr(cat), T * B, H, W, hid, cat.shape[-1], *self.ks,
                                                 _lib.stream_ptr())
            _lib.check(rc, "eas_spp_pool_fwd")
        else:
            y4 = y.flatten(0, 1).permute(0, 3, 1, 2)   # [T*B, C, H, W] view, channels-last memory
            for i, k in enumerate(self.ks):
                p = F.max_pool2d(y4, k, 1, k // 2)
                cat[..., (i + 1) * hid:(i + 2) * hid].copy_(p.permute(0, 2, 3, 1).reshape(T, B, H, W, hid))
        return self.conv2.run(cat, T)

    def forward(self, x):
        x = self.conv1(x)
        # (training path) PyTorch's channels-last max-pool kernels are ~5x slower than its NCHW ones at these window
        # sizes -- 0.81 of a 6.9 ms SYOLOX-S step in profiles/r2_launches_train_summary.txt: pool an NCHW copy
        xp = x.contiguous() if not x.is_contiguous() else x
        return self.conv2(_cat_channels([x] + [m(xp) for m in self.m]))


def _cat_channels(parts):
    """``torch.cat(parts, dim=-3)`` for ``[T, B, C, H, W]`` activations (training path).  When the parts are views of
    channels-last buffers (cuDNN NHWC kernels, the neurons working in place) the concatenation is done on the memory as it
    lies and handed back as the same kind of view; ``torch.cat`` would produce an NCHW tensor there and the next
    channels-last convolution would convert it back (one copy kernel per concatenation and another per consumer)."""
    if all(p.dim() == 5 and not p.is_contiguous() and p.permute(0, 1, 3, 4, 2).is_contiguous() for p in parts):
        return torch.cat([p.permute(0, 1, 3, 4, 2) for p in parts], dim=-1).permute(0, 1, 4, 2, 3)
    return torch.cat(parts, dim=-3)


class _CSPLayer(nn.Module):
    def __init__(self, cin, cout, n, shortcut, spike_fn):
        super().__init__()
        hid = int(cout * 0.5)
        self.conv1 = FusedConvBNPLIF(cin, hid, 1, 1, spike_fn)
        self.conv2 = FusedConvBNPLIF(cin, hid, 1, 1, spike_fn)
        self.conv3 = FusedConvBNPLIF(2 * hid, cout, 1, 1, spike_fn)
        self.m = nn.Sequential(*[_Bottleneck(hid, hid, shortcut, 1.0, spike_fn) for _ in range(n)])

    def run(self, x, T, out=None):
        Tn, B, H, W, _ = x.shape
        hid = self.conv1.conv[0].out_channels
        cat = torch.empty((T, B, H, W, 2 * hid), dtype=ACT_DTYPE, device=x.device)
        self.conv2.run(x, T, out=cat[..., hid:])
        blocks = list(self.m)
        y = self.conv1.run(x, T, out=None if blocks else cat[..., :hid])
        for i, blk in enumerate(blocks):
            y = blk.run(y, T, out=cat[..., :hid] if i == len(blocks) - 1 else None)
        return self.conv3.run(cat, T, out=out)

    def forward(self, x):
        return self.conv3(_cat_channels([self.m(self.conv1(x)), self.conv2(x)]))


class SpikingCSPDarknet(nn.Module):
    """Spiking CSPDarknet (darknet.py:97-180 after convert_to_spiking), inference, channels-last.

    ``forward(frames)``: ``frames`` is the sampler output ``[Ts or T, B, 2, H, W]`` fp32; a single
    frame (Ts == 1) is broadcast over the T SNN steps as ``SpikingYOLOX.forward`` does
    (spiking_yolox.py:54-55) -- here without materialising the copies: the stem and the first
    spiking conv are computed once and only the neuron runs T times.
    Returns ``{dark3, dark4, dark5}`` spike tensors as logical ``[T, B, C, H, W]`` views.
    """

    def __init__(self, dep_mul, wid_mul, in_dim=2, spike_fn=None, T=3, out_features=("dark3", "dark4", "dark5")):
        super().__init__()
        spike_fn = spike_fn if spike_fn is not None else ATan(2.0)
        c = int(wid_mul * 64)
        d = max(round(dep_mul * 3), 1)
        self.T = T
        self.out_features = out_features
        self.stem = SeqToANNContainer(_Focus(in_dim, c, 3))
        self.dark2 = nn.Sequential(FusedConvBNPLIF(c, c * 2, 3, 2, spike_fn), _CSPLayer(c * 2, c * 2, d, True, spike_fn))
        self.dark3 = nn.Sequential(FusedConvBNPLIF(c * 2, c * 4, 3, 2, spike_fn),
                                   _CSPLayer(c * 4, c * 4, d * 3, True, spike_fn))
        self.dark4 = nn.Sequential(FusedConvBNPLIF(c * 4, c * 8, 3, 2, spike_fn),
                                   _CSPLayer(c * 8, c * 8, d * 3, True, spike_fn))
        self.dark5 = nn.Sequential(FusedConvBNPLIF(c * 8, c * 16, 3, 2, spike_fn), _SPP(c * 16, c * 16, spike_fn),
                                   _CSPLayer(c * 16, c * 16, d, False, spike_fn))

    @torch.no_grad()
    def run_cl(self, frames: torch.Tensor, into: dict | None = None) -> dict:
        """All stage outputs as channels-last fp16 spike tensors ``[T, B, H, W, C]`` (what the fused FPN reads).
        ``into[name]`` (optional) = a function ``(T, B, H, W, C) -> tensor`` giving the buffer (e.g. a channel slice of
        a concatenation buffer of the pyramid) that stage ``name`` writes its output into."""
        if self.training:
            raise RuntimeError("run_cl is the fused inference path; call .eval() (training: forward())")
        _lib.require_cuda(frames)
        T = self.T
        if frames.shape[0] not in (1, T):
            raise ValueError("the timestep of SNN is not matched with that of input")   # spiking_yolox.py:57
        stem2 = self.stem[0].run(frames.float())                     # [2, Tx, B, H/2, W/2, c] fp16 planes
        x = self.dark2[0].run(stem2, T, n_xsplit=2)                   # first spiking conv: real-valued input
        outs = {}
        x = self.dark2[1].run(x, T)
        outs["dark2"] = x
        for name in ("dark3", "dark4", "dark5"):
            seq = getattr(self, name)
            x = seq[0].run(x, T)
            blocks = list(seq)[1:]
            for i, blk in enumerate(blocks):
                if i == len(blocks) - 1 and into is not None and name in into:
                    _, B, H, W, _ = x.shape
                    x = blk.run(x, T, out=into[name](T, B, H, W, blk.conv3.conv[0].out_channels))
                else:
                    x = blk.run(x, T)
            outs[name] = x
        return outs

    def forward_train(self, frames: torch.Tensor, return_all: bool = False):
        """Training path (cfg 4): batch-statistics BN and autograd need the conv / BN outputs, so those two run as
        PyTorch ops; every neuron (forward and surrogate-gradient backward) runs on ``eas_plif_fwd / eas_plif_bwd``.
        ``frames`` ``[Ts or T, B, 2, H, W]``; the Ts == 1 frame is broadcast as spiking_yolox.py:54-55 does.
        With the module converted by ``.to(memory_format=torch.channels_last)`` cuDNN runs NHWC kernels end to end and
        the neurons take the channels-last views as they are (SYOLOX-S, 8 windows: 9.6 -> 7.3 ms per replayed step)."""
        _lib.require_cuda(frames)
        T = self.T
        if frames.shape[0] == 1:
            frames = frames.expand(T, -1, -1, -1, -1)
        if frames.shape[0] != T:
            raise ValueError("the timestep of SNN is not matched with that of input")
        outs = {}
        bump_bn_counters(self)                       # every layer's num_batches_tracked, one kernel
        x = self.stem(frames)
        for name in ("dark2", "dark3", "dark4", "dark5"):
            x = getattr(self, name)(x)
            outs[name] = x
        keys = outs.keys() if return_all else self.out_features
        return {k: outs[k] for k in keys}

    def forward(self, frames: torch.Tensor, return_all: bool = False):
        if self.training:
            return self.forward_train(frames, return_all)
        return self._forward_eval(frames, return_all)

    @torch.no_grad()
    def _forward_eval(self, frames: torch.Tensor, return_all: bool = False):
        outs = self.run_cl(frames)
        keys = outs.keys() if return_all else self.out_features
        return {k: outs[k].permute(0, 1, 4, 2, 3) for k in keys}


class GraphedForward:
    """CUDA-graph replay of a fused inference forward (``SpikingCSPDarknet`` or any module of fused layers).

    The ~60 launches of a SYOLOX backbone forward (51 conv+BN+PLIF kernels + glue) are launch bound for small
    batches; captured once into a ``torch.cuda.CUDAGraph`` they replay with one host call.  Input and outputs
    live in static buffers: ``__call__`` copies the new frames in and returns the (reused) output tensors."""

    def __init__(self, module: nn.Module, example: torch.Tensor, warmup: int = 2):
        _lib.require_cuda(example)
        self.module = module
        self.static_in = example.clone()
        side = torch.cuda.Stream(example.device)
        side.wait_stream(torch.cuda.current_stream(example.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):                      # packs weights, sizes the allocator pools
                module(self.static_in)
        torch.cuda.current_stream(example.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = module(self.static_in)

    def __call__(self, frames: torch.Tensor):
        if frames.shape != self.static_in.shape or frames.dtype != self.static_in.dtype:
            raise ValueError("GraphedForward was captured for %s %s" % (tuple(self.static_in.shape), self.static_in.dtype))
        self.static_in.copy_(frames)
        self.graph.replay()
        return self.static_out


class GraphedTrainStep:
    """CUDA-graph replay of one training step (the reference's ``train_one_iter``, yolox/core/trainer.py:96-125:
    forward, loss, backward, optimizer step, ``reset_net``) for launch-bound sizes: a SYOLOX-S step on 8 windows is
    ~1500 kernel launches (cuDNN conv / BN, the PLIF forward / backward kernels, the sampler's BPTT, element-wise glue,
    Adam) that the host issues in 16 ms while the GPU needs a third of that.

    Two graphs around the one collective: ``backward`` graph = zero the gradients, ``loss_fn(*static inputs)``,
    ``loss.backward()``; then the gradient all-reduce runs eagerly (``allreduce``: ``"flat"`` = one NCCL all-reduce of the
    flat gradient buffer, or a callable taking the parameters, or None); then the
    ``update`` graph = ``optimizer.step()`` (the optimizer must be built with ``capturable=True``) and whatever
    ``after`` does on the device.  Python-side state (the neurons' ``v`` handles) is left as ``reset_net`` leaves it.

    ``loss_fn`` takes the static input tensors and returns the scalar loss; ``__call__(*inputs)`` copies new inputs
    into the static buffers, replays, and returns the (static) loss tensor.  Build it before any eager backward through
    the same parameters, or after every tensor of those eager graphs has been dropped: a live graph keeps the
    parameters' gradient accumulators bound to the stream they first ran on, which a capture may not wait for."""

    def __init__(self, loss_fn, inputs, params, optimizer, allreduce=None, after=None, warmup: int = 3):
        params = list(params)
        dev = params[0].device
        self.static_in = [t.clone() for t in inputs]
        self.allreduce, self.params = allreduce, params
        # Every gradient is a view (with the parameter's own strides) of ONE flat buffer: zeroing the gradients is one
        # kernel instead of one per parameter (147 of them for SYOLOX-S, ~2 us each under replay), and the all-reduce of
        # a data-parallel run (allreduce="flat") is one collective on that buffer with no pack / unpack copies.
        self.flat_grad = torch.zeros(sum(q.numel() for q in params), dtype=params[0].dtype, device=dev)
        o = 0
        for q in params:
            dense = q.is_contiguous() or q.is_contiguous(memory_format=torch.channels_last) if q.dim() == 4 \
                else q.is_contiguous()
            q.grad = (self.flat_grad[o:o + q.numel()].as_strided(q.size(), q.stride())
                      if dense and q.dtype == self.flat_grad.dtype and q.device == dev else torch.zeros_like(q))
            o += q.numel()
        self._loose = [q.grad for q in params if q.grad.untyped_storage().data_ptr() != self.flat_grad.untyped_storage().data_ptr()]

        def fwd_bwd():
            self.flat_grad.zero_()                   # (grads stay allocated: the update graph reads these tensors)
            for g in self._loose:
                g.zero_()
            loss = loss_fn(*self.static_in)
            loss.backward()
            return loss.detach()

        def reduce_grads():
            if allreduce == "flat":
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                    dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)
                    self.flat_grad.div_(dist.get_world_size())
                    for g in self._loose:
                        dist.all_reduce(g, op=dist.ReduceOp.SUM)
                        g.div_(dist.get_world_size())
            elif allreduce is not None:
                allreduce(params)

        self._reduce_grads = reduce_grads

        def update():
            optimizer.step()
            if after is not None:
                after()

        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):                  # cuDNN picks its algorithms, gradients and Adam state get allocated
                fwd_bwd()
                reduce_grads()
                update()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.g_bwd, self.g_upd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_bwd):
            self.loss = fwd_bwd()
        pool = self.g_bwd.pool()
        with torch.cuda.graph(self.g_upd, pool=pool):
            update()

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            if src is not dst:
                dst.copy_(src)
        self.g_bwd.replay()
        self._reduce_grads()
        self.g_upd.replay()
        return self.loss


def convert_to_spiking(model: nn.Module, spike_fn, fuse: bool = True, n_wsplit: int = 2) -> nn.Module:
    """``yolox/utils/utils_snn.py:16-58`` with the fused layer: every child that looks like the
    reference's ``BaseConv`` (``.conv`` Conv2d, ``.bn`` BatchNorm2d, ``.act``) becomes a
    :class:`FusedConvBNPLIF` (same keys); ``Focus`` is wrapped whole and stays ANN; lone Conv2d /
    Upsample / MaxPool2d are wrapped; remaining activations become :class:`ParametricLIFNode`."""
    for name, module in model.named_children():
        cname = type(module).__name__
        if cname == "Focus":
            setattr(model, name, SeqToANNContainer(module))
        elif fuse and isinstance(getattr(module, "conv", None), nn.Conv2d) and \
                isinstance(getattr(module, "bn", None), nn.BatchNorm2d) and hasattr(module, "act") and \
                module.conv.groups == 1:
            setattr(model, name, FusedConvBNPLIF.from_modules(module.conv, module.bn, spike_fn=spike_fn,
                                                              n_wsplit=n_wsplit))
        elif isinstance(module, (nn.Conv2d, nn.Upsample, nn.MaxPool2d)):
            setattr(model, name, SeqToANNContainer(module))
        elif isinstance(module, nn.BatchNorm2d):
            bn = MultiStepBatchNorm2d(module.num_features, module.eps, module.momentum)
            setattr(model, name, bn)
        elif name.endswith("act") or isinstance(module, (nn.ReLU, nn.SiLU, nn.LeakyReLU)):
            setattr(model, name, ParametricLIFNode(init_tau=2.0, decay_input=False, v_threshold=1.0, v_reset=None,
                                                   surrogate_function=copy.deepcopy(spike_fn), detach_reset=False,
                                                   step_mode="m", backend="torch"))
        else:
            convert_to_spiking(module, spike_fn, fuse=fuse, n_wsplit=n_wsplit)
    return model
