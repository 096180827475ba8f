"""(f-2) The whole SYOLOX detector, inference: sampler -> spiking CSPDarknet -> ANN PAFPN -> YOLOX head -> decode.

Mirrors the model the reference builds for ``use_spike True`` (``yolox/exp/event_yolox_base.py:188-203``):
``SpikingYOLOX(SpikingYOLOPAFPN(...), YOLOXHead(...), embedding, T)`` with identical child names, so a reference
checkpoint's ``state_dict`` loads with ``strict=True``:

* :class:`SpikingYOLOPAFPN` -- ``yolox/models/spiking_yolo_pafpn.py:14-120``: spiking backbone, firing rate over the
  T steps (``.mean(axis=0)``, :98), then the ANN top-down / bottom-up pyramid (conv -> BN -> SiLU);
* :class:`YOLOXHead`        -- ``yolox/models/yolo_head.py:17-250`` (inference branch: stems, cls / reg towers,
  1x1 predictors, sigmoid, flatten + concat over levels, ``decode_outputs``);
* :class:`SpikingYOLOX`     -- ``yolox/models/spiking_yolox.py:23-74``;
* :func:`postprocess`       -- ``yolox/utils/boxes.py:33-77`` (confidence filter + class-wise NMS).

Every convolution runs on the tensor-core kernel ``eas_conv_bn_plif_fwd`` (folded BN, fp16 hi/lo split weights
and activations = fp32-equivalent products, SiLU or raw output in the epilogue); real-valued activations are kept
as two fp16 planes ``[2, 1, B, H, W, C]`` channels-last; the concatenations of the pyramid are channel slices of one
buffer that producers write into directly.  Training of these ANN parts is out of scope (inference path).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _lib
from .fused import (ACT_DTYPE, OUT_PREACT, AnnBaseConv, FusedConvBNPLIF, SeqToANNContainer, SpikingCSPDarknet, _CSPLayer,
                    conv_bn_plif, pack_weight)


# ------------------------------------------------------------------------------------------------
# glue kernels
# ------------------------------------------------------------------------------------------------
def _planes_ld(p: torch.Tensor) -> int:
    """pixel stride (elements) of a planes tensor ``[2, 1, B, H, W, C]`` that may be a channel slice."""
    _, _, B, H, W, C = p.shape
    ld = p.stride(-2) if W > 1 else (p.stride(-3) if H > 1 else (p.stride(-4) if B > 1 else C))
    want = (None, B * H * W * ld, H * W * ld, W * ld, ld, 1)
    if p.dtype != ACT_DTYPE or ld < C or any(d > 1 and s != w for s, w, d in list(zip(p.stride(), want, p.shape))[2:]):
        raise ValueError("expected fp16 channels-last planes [2, 1, B, H, W, C] (or a channel slice of such a buffer)")
    return ld


def time_mean_planes(spikes: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """``spikes.mean(axis=0)`` (spiking_yolo_pafpn.py:98) of channels-last fp16 spikes ``[T, B, H, W, C]`` into the
    planes (slice) ``out [2, 1, B, H, W, C]``."""
    _lib.require_cuda(spikes, out)
    T, B, H, W, C = spikes.shape
    if spikes.dtype != ACT_DTYPE or not spikes.is_contiguous() or tuple(out.shape) != (2, 1, B, H, W, C):
        raise ValueError("time_mean_planes: spikes [T,B,H,W,C] fp16 contiguous, out [2,1,B,H,W,C]")
    ld = _planes_ld(out)
    with torch.cuda.device(spikes.device):
        rc = _lib.lib().eas_time_mean_planes(_lib.ptr(spikes), T, B * H * W, C, C, _lib.ptr(out), ld, out.stride(0),
                                             _lib.stream_ptr())
    _lib.check(rc, "eas_time_mean_planes")
    return out


def upsample2x_planes(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """``nn.Upsample(scale_factor=2, mode='nearest')`` on planes ``[2, 1, B, H, W, C]`` -> ``[2, 1, B, 2H, 2W, C]``."""
    _lib.require_cuda(x, out)
    _, _, B, H, W, C = x.shape
    if tuple(out.shape) != (2, 1, B, 2 * H, 2 * W, C):
        raise ValueError("upsample2x_planes: out must be [2, 1, B, 2H, 2W, C]")
    with torch.cuda.device(x.device):
        rc = _lib.lib().eas_upsample2x_planes(_lib.ptr(x), 2, x.stride(0), B, H, W, C, _planes_ld(x), _lib.ptr(out),
                                              _planes_ld(out), out.stride(0), _lib.stream_ptr())
    _lib.check(rc, "eas_upsample2x_planes")
    return out


def _new_planes(B, H, W, C, device):
    return torch.empty((2, 1, B, H, W, C), dtype=ACT_DTYPE, device=device)


# ------------------------------------------------------------------------------------------------
# ANN blocks on planes (network_blocks.py:31-56, 81-104, 150-188 with act = SiLU)
# ------------------------------------------------------------------------------------------------
class _AnnBottleneck(nn.Module):
    def __init__(self, cin, cout, shortcut, expansion=0.5):
        super().__init__()
        if shortcut and cin == cout:
            raise NotImplementedError("the PAFPN builds its CSP layers with shortcut=False")
        hid = int(cout * expansion)
        self.conv1 = AnnBaseConv(cin, hid, 1, 1)
        self.conv2 = AnnBaseConv(hid, cout, 3, 1)

    def run(self, xp, out=None):
        return self.conv2.run(self.conv1.run(xp), out=out)


class _AnnCSPLayer(nn.Module):
    """``CSPLayer`` (network_blocks.py:150-188): conv3(cat(m(conv1(x)), conv2(x)))."""

    def __init__(self, cin, cout, n, shortcut):
        super().__init__()
        hid = int(cout * 0.5)
        self.conv1 = AnnBaseConv(cin, hid, 1, 1)
        self.conv2 = AnnBaseConv(cin, hid, 1, 1)
        self.conv3 = AnnBaseConv(2 * hid, cout, 1, 1)
        self.m = nn.Sequential(*[_AnnBottleneck(hid, hid, shortcut, 1.0) for _ in range(n)])

    def run(self, xp, out=None):
        _, _, B, H, W, _ = xp.shape
        hid = self.conv1.conv.out_channels
        cat = _new_planes(B, H, W, 2 * hid, xp.device)
        self.conv2.run(xp, out=cat[..., hid:])
        blocks = list(self.m)
        y = self.conv1.run(xp, out=None if blocks else cat[..., :hid])
        for i, blk in enumerate(blocks):
            y = blk.run(y, out=cat[..., :hid] if i == len(blocks) - 1 else None)
        return self.conv3.run(cat, out=out)


class SpikingYOLOPAFPN(nn.Module):
    """``SpikingYOLOPAFPN`` (spiking_yolo_pafpn.py:14-120), inference.  ``forward(frames)`` takes the sampler
    output ``[Ts or T, B, in_dim, H, W]`` and returns ``(pan_out2, pan_out1, pan_out0)`` as fp32
    ``[B, C, H, W]`` tensors; :meth:`run` keeps them as planes for the fused head."""

    def __init__(self, depth=1.0, width=1.0, in_features=("dark3", "dark4", "dark5"), in_channels=(256, 512, 1024),
                 in_dim=2, spike_fn=None, T=3):
        super().__init__()
        self.backbone = SpikingCSPDarknet(depth, width, in_dim=in_dim, spike_fn=spike_fn, T=T,
                                          out_features=tuple(in_features))
        self.in_features = tuple(in_features)
        c0, c1, c2 = (int(c * width) for c in in_channels)
        n = round(3 * depth)
        self.upsample = nn.Upsample(scale_factor=2, mode="nearest")
        self.lateral_conv0 = AnnBaseConv(c2, c1, 1, 1)
        self.C3_p4 = _AnnCSPLayer(2 * c1, c1, n, False)
        self.reduce_conv1 = AnnBaseConv(c1, c0, 1, 1)
        self.C3_p3 = _AnnCSPLayer(2 * c0, c0, n, False)
        self.bu_conv2 = AnnBaseConv(c0, c0, 3, 2)
        self.C3_n3 = _AnnCSPLayer(2 * c0, c1, n, False)
        self.bu_conv1 = AnnBaseConv(c1, c1, 3, 2)
        self.C3_n4 = _AnnCSPLayer(2 * c1, c2, n, False)
        self.channels = (c0, c1, c2)

    @torch.no_grad()
    def run(self, frames: torch.Tensor):
        if self.training:
            raise RuntimeError("SpikingYOLOPAFPN (fused) is the inference path; call .eval()")
        feats = self.backbone.run_cl(frames)
        s2, s1, s0 = (feats[f] for f in self.in_features)          # spikes [T, B, H, W, C] at strides 8 / 16 / 32
        c0, c1, c2 = self.channels
        dev = frames.device
        _, B, H0, W0, _ = s0.shape
        _, _, H1, W1, _ = s1.shape
        _, _, H2, W2, _ = s2.shape
        if (H1, W1) != (2 * H0, 2 * W0) or (H2, W2) != (2 * H1, 2 * W1):
            raise ValueError("input height and width must be multiples of 32 (event_yolox_base.py:556-559)")
        x0 = time_mean_planes(s0, _new_planes(B, H0, W0, c2, dev))
        cat_p4 = _new_planes(B, H1, W1, 2 * c1, dev)               # [upsample(fpn_out0), x1]      (:102-103)
        cat_p3 = _new_planes(B, H2, W2, 2 * c0, dev)               # [upsample(fpn_out1), x2]      (:107-108)
        cat_n3 = _new_planes(B, H1, W1, 2 * c0, dev)               # [bu_conv2(pan_out2), fpn_out1] (:111-112)
        cat_n4 = _new_planes(B, H0, W0, 2 * c1, dev)               # [bu_conv1(pan_out1), fpn_out0] (:115-116)
        time_mean_planes(s1, cat_p4[..., c1:])
        time_mean_planes(s2, cat_p3[..., c0:])
        fpn_out0 = self.lateral_conv0.run(x0, out=cat_n4[..., c1:])
        upsample2x_planes(fpn_out0, cat_p4[..., :c1])
        f_out0 = self.C3_p4.run(cat_p4)
        fpn_out1 = self.reduce_conv1.run(f_out0, out=cat_n3[..., c0:])
        upsample2x_planes(fpn_out1, cat_p3[..., :c0])
        pan_out2 = self.C3_p3.run(cat_p3)
        self.bu_conv2.run(pan_out2, out=cat_n3[..., :c0])
        pan_out1 = self.C3_n3.run(cat_n3)
        self.bu_conv1.run(pan_out1, out=cat_n4[..., :c1])
        pan_out0 = self.C3_n4.run(cat_n4)
        return pan_out2, pan_out1, pan_out0

    def forward(self, frames: torch.Tensor):
        return tuple(planes_to_nchw(p) for p in self.run(frames))


def upsample2x_spikes(x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """``SeqToANNContainer(nn.Upsample(scale_factor=2, mode='nearest'))`` (utils_snn.py:25-27) on channels-last
    fp16 spikes ``[T, B, H, W, C]`` into ``out [T, B, 2H, 2W, C]`` (a channel slice of a concatenation buffer)."""
    _lib.require_cuda(x, out)
    T, B, H, W, C = x.shape
    if tuple(out.shape) != (T, B, 2 * H, 2 * W, C) or x.dtype != ACT_DTYPE or out.dtype != ACT_DTYPE or not x.is_contiguous():
        raise ValueError("upsample2x_spikes: x [T,B,H,W,C] fp16 contiguous, out [T,B,2H,2W,C]")
    ld = out.stride(-2)
    if out.stride(-1) != 1 or out.stride(-3) != 2 * W * ld or out.stride(1) != 4 * H * W * ld or \
            (T > 1 and out.stride(0) != B * 4 * H * W * ld):
        raise ValueError("upsample2x_spikes: out must be a channel slice of a contiguous [T,B,2H,2W,C'] buffer")
    with torch.cuda.device(x.device):   # (T, B) is one batch axis for the kernel: one plane
        rc = _lib.lib().eas_upsample2x_planes(_lib.ptr(x), 1, 0, T * B, H, W, C, C, _lib.ptr(out), ld, 0,
                                              _lib.stream_ptr())
    _lib.check(rc, "eas_upsample2x_planes")
    return out


class FullSpikeYOLOPAFPN(nn.Module):
    """``convert_to_spiking(YOLOPAFPN(...))`` -- the backbone of ``use_spike full_spike / full_spike_v2``
    (event_yolox_base.py:207-208; yolo_pafpn.py:16-116 through utils_snn.py:16-58): the spiking CSPDarknet AND a
    spiking pyramid.  Every ``BaseConv`` is conv -> BN -> PLIF over the T steps on spike inputs (one fp16 plane, two
    weight planes: fewer product terms than the ANN pyramid and the LIF in the conv epilogue), the upsampling and the
    concatenations (``torch.cat(.., -3)``, :103-116) act per time step.  Returns the three spike tensors
    ``(pan_out2, pan_out1, pan_out0)``; reference child names, so its ``state_dict`` loads with ``strict=True``."""

    def __init__(self, depth=1.0, width=1.0, in_features=("dark3", "dark4", "dark5"), in_channels=(256, 512, 1024),
                 in_dim=2, spike_fn=None, T=3):
        super().__init__()
        self.backbone = SpikingCSPDarknet(depth, width, in_dim=in_dim, spike_fn=spike_fn, T=T,
                                          out_features=tuple(in_features))
        self.in_features = tuple(in_features)
        c0, c1, c2 = (int(c * width) for c in in_channels)
        n = round(3 * depth)
        self.upsample = SeqToANNContainer(nn.Upsample(scale_factor=2, mode="nearest"))
        self.lateral_conv0 = FusedConvBNPLIF(c2, c1, 1, 1, spike_fn)
        self.C3_p4 = _CSPLayer(2 * c1, c1, n, False, spike_fn)
        self.reduce_conv1 = FusedConvBNPLIF(c1, c0, 1, 1, spike_fn)
        self.C3_p3 = _CSPLayer(2 * c0, c0, n, False, spike_fn)
        self.bu_conv2 = FusedConvBNPLIF(c0, c0, 3, 2, spike_fn)
        self.C3_n3 = _CSPLayer(2 * c0, c1, n, False, spike_fn)
        self.bu_conv1 = FusedConvBNPLIF(c1, c1, 3, 2, spike_fn)
        self.C3_n4 = _CSPLayer(2 * c1, c2, n, False, spike_fn)
        self.channels = (c0, c1, c2)

    @torch.no_grad()
    def run(self, frames: torch.Tensor):
        """frames ``[Ts or T, B, in_dim, H, W]`` -> spikes ``[T, B, H/s, W/s, C]`` (fp16, channels-last) at s = 8, 16, 32."""
        if self.training:
            raise RuntimeError("FullSpikeYOLOPAFPN (fused) is the inference path; call .eval()")
        T = self.backbone.T
        c0, c1, c2 = self.channels
        dev = frames.device
        cats = {}

        def buf(name, T_, B, H, W, ctot):
            cats[name] = torch.empty((T_, B, H, W, ctot), dtype=ACT_DTYPE, device=dev)
            return cats[name]

        # dark4 / dark3 write straight into the second halves of the top-down concatenations (:102-108)
        feats = self.backbone.run_cl(frames, into={
            "dark4": lambda T_, B, H, W, C: buf("p4", T_, B, H, W, 2 * c1)[..., c1:],
            "dark3": lambda T_, B, H, W, C: buf("p3", T_, B, H, W, 2 * c0)[..., c0:]})
        s0 = feats[self.in_features[2]]
        _, B, H0, W0, _ = s0.shape
        cat_p4, cat_p3 = cats["p4"], cats["p3"]
        H1, W1, H2, W2 = cat_p4.shape[2], cat_p4.shape[3], cat_p3.shape[2], cat_p3.shape[3]
        if (H1, W1) != (2 * H0, 2 * W0) or (H2, W2) != (2 * H1, 2 * W1):
            raise ValueError("input height and width must be multiples of 32 (event_yolox_base.py:556-559)")
        cat_n3 = torch.empty((T, B, H1, W1, 2 * c0), dtype=ACT_DTYPE, device=dev)    # [bu_conv2(pan_out2), fpn_out1]
        cat_n4 = torch.empty((T, B, H0, W0, 2 * c1), dtype=ACT_DTYPE, device=dev)    # [bu_conv1(pan_out1), fpn_out0]
        fpn_out0 = self.lateral_conv0.run(s0, T, out=cat_n4[..., c1:])
        upsample2x_spikes(fpn_out0.contiguous() if not fpn_out0.is_contiguous() else fpn_out0, cat_p4[..., :c1])
        f_out0 = self.C3_p4.run(cat_p4, T)
        fpn_out1 = self.reduce_conv1.run(f_out0, T, out=cat_n3[..., c0:])
        upsample2x_spikes(fpn_out1.contiguous() if not fpn_out1.is_contiguous() else fpn_out1, cat_p3[..., :c0])
        pan_out2 = self.C3_p3.run(cat_p3, T)
        self.bu_conv2.run(pan_out2, T, out=cat_n3[..., :c0])
        pan_out1 = self.C3_n3.run(cat_n3, T)
        self.bu_conv1.run(pan_out1, T, out=cat_n4[..., :c1])
        pan_out0 = self.C3_n4.run(cat_n4, T)
        return pan_out2, pan_out1, pan_out0

    def forward(self, frames: torch.Tensor):
        """Reference-shaped: spike tensors as logical ``[T, B, C, H, W]`` views."""
        return tuple(p.permute(0, 1, 4, 2, 3) for p in self.run(frames))


def planes_to_nchw(p: torch.Tensor) -> torch.Tensor:
    """planes ``[2, 1, B, H, W, C]`` -> fp32 ``[B, C, H, W]`` (hi + lo)."""
    return (p[0, 0].float() + p[1, 0].float()).permute(0, 3, 1, 2)


# ------------------------------------------------------------------------------------------------
# head
# ------------------------------------------------------------------------------------------------
class YOLOXHead(nn.Module):
    """``YOLOXHead`` (yolo_head.py:17-250), inference branch."""

    def __init__(self, num_classes, width=1.0, strides=(8, 16, 32), in_channels=(256, 512, 1024)):
        super().__init__()
        self.num_classes = num_classes
        self.decode_in_inference = True
        self.strides = list(strides)
        hid = int(256 * width)
        self.stems = nn.ModuleList(AnnBaseConv(int(c * width), hid, 1, 1) for c in in_channels)
        self.cls_convs = nn.ModuleList(nn.Sequential(AnnBaseConv(hid, hid, 3, 1), AnnBaseConv(hid, hid, 3, 1))
                                       for _ in in_channels)
        self.reg_convs = nn.ModuleList(nn.Sequential(AnnBaseConv(hid, hid, 3, 1), AnnBaseConv(hid, hid, 3, 1))
                                       for _ in in_channels)
        self.cls_preds = nn.ModuleList(nn.Conv2d(hid, num_classes, 1, 1, 0) for _ in in_channels)
        self.reg_preds = nn.ModuleList(nn.Conv2d(hid, 4, 1, 1, 0) for _ in in_channels)
        self.obj_preds = nn.ModuleList(nn.Conv2d(hid, 1, 1, 1, 0) for _ in in_channels)
        self._pred_cache = {}
        self.fp16_inputs = False

    def initialize_biases(self, prior_prob):                    # yolo_head.py:130-140
        for k in range(len(self.cls_preds)):
            for name in ("cls_preds", "obj_preds"):
                self._pred(name, k).bias.data.fill_(-math.log((1 - prior_prob) / prior_prob))

    def _packed_preds(self, k):
        """(reg | obj) as one 5-channel 1x1 conv (both read reg_feat), cls as another; no BN, conv bias kept."""
        reg, obj, cls = (self._pred(n, k) for n in ("reg_preds", "obj_preds", "cls_preds"))
        src = [reg.weight, obj.weight, cls.weight, reg.bias, obj.bias, cls.bias]
        key = tuple((t.data_ptr(), t._version) for t in src)
        hit = self._pred_cache.get(k)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                w_ro = torch.cat((reg.weight, obj.weight), 0)
                b_ro = torch.cat((reg.bias, obj.bias), 0).float().contiguous()
                hit = (key, pack_weight(w_ro, 2) + (b_ro,), pack_weight(cls.weight, 2) + (cls.bias.float().contiguous(),))
                self._pred_cache[k] = hit
        return hit[1], hit[2]

    def _pred(self, name, k) -> nn.Conv2d:
        return getattr(self, name)[k]

    @torch.no_grad()
    def run(self, feats):
        """feats: planes ``[2, 1, B, H, W, C]`` per level -> ``[B, n_anchors, 5 + num_classes]`` fp32."""
        if self.training:
            raise RuntimeError("YOLOXHead (fused) is the inference path; call .eval()")
        towers = []
        for k, xp in enumerate(feats):
            x = self.stems[k].run(xp)
            towers.append((self.cls_convs[k][1].run(self.cls_convs[k][0].run(x)),
                           self.reg_convs[k][1].run(self.reg_convs[k][0].run(x))))
        return self._predict_and_decode(towers)

    @torch.no_grad()
    def _predict_and_decode(self, towers):
        """(cls_feat, reg_feat) planes per level -> 1x1 predictors -> sigmoid / grid decode -> ``[B, A, 5 + classes]``."""
        n_ch = 5 + self.num_classes
        B = towers[0][0].shape[2]
        dev = towers[0][0].device
        hw = [tuple(c.shape[3:5]) for c, _ in towers]
        self.hw = hw
        A = sum(h * w for h, w in hw)
        out = torch.empty((B, A, n_ch), dtype=torch.float32, device=dev)
        a_off = 0
        L = _lib.lib()
        for k, (cls_feat, reg_feat) in enumerate(towers):
            H, W = hw[k]
            (w_ro, u_ro, b_ro), (w_c, u_c, b_c) = self._packed_preds(k)
            preds = torch.empty((1, B, H, W, n_ch), dtype=torch.float32, device=dev)
            nx = 1 if self.fp16_inputs else 2
            conv_bn_plif(reg_feat[:nx], w_ro, b_ro, None, 1, 1, 1, n_xsplit=nx, out=preds[..., :5],
                         out_mode=OUT_PREACT, w_unscale=u_ro)
            conv_bn_plif(cls_feat[:nx], w_c, b_c, None, 1, 1, 1, n_xsplit=nx, out=preds[..., 5:],
                         out_mode=OUT_PREACT, w_unscale=u_c)
            with torch.cuda.device(dev):
                rc = L.eas_yolox_decode(_lib.ptr(preds), B, H, W, n_ch, n_ch, float(self.strides[k]),
                                        int(self.decode_in_inference), _lib.ptr(out), a_off, A, _lib.stream_ptr())
            _lib.check(rc, "eas_yolox_decode")
            a_off += H * W
        return out

    def forward(self, xin, labels=None, imgs=None):
        """Reference-shaped entry: ``xin`` = fp32 ``[B, C, H, W]`` per level."""
        return self.run([nchw_to_planes(x) for x in xin])


class SpikingYOLOXHead(YOLOXHead):
    """``SpikingYOLOXHead`` (spiking_yolo_head.py:18-230), inference branch, on the spike tensors of
    :class:`FullSpikeYOLOPAFPN`.

    ``full_spike=False`` (``use_spike full_spike``): the head itself stays ANN; every level is averaged over the T
    steps first (``x.mean(axis=0)``, :159-160) -- :class:`YOLOXHead` on firing rates.
    ``full_spike=True`` (``full_spike_v2``): stems and towers are conv -> BN -> PLIF over the T steps
    (``convert_to_spiking(self)``, :125-127), the 1x1 predictors run per step and their outputs are averaged over T
    (:173-178).  The predictors are linear, so the average is taken on the tower spikes instead (firing rates k/T, then
    ONE predictor pass): the same number up to fp32 rounding, a third of the work.  State-dict keys as the reference's
    (``stems.0.conv.0.weight``, ``stems.0.act.w``, ``cls_preds.0.0.weight``)."""

    def __init__(self, num_classes, width=1.0, strides=(8, 16, 32), in_channels=(256, 512, 1024), spike_fn=None,
                 full_spike=False, T=3):
        super().__init__(num_classes, width, strides, in_channels)
        self.full_spike = bool(full_spike)
        self.T = T
        if self.full_spike:
            hid = int(256 * width)
            self.stems = nn.ModuleList(FusedConvBNPLIF(int(c * width), hid, 1, 1, spike_fn) for c in in_channels)
            self.cls_convs = nn.ModuleList(nn.Sequential(FusedConvBNPLIF(hid, hid, 3, 1, spike_fn),
                                                         FusedConvBNPLIF(hid, hid, 3, 1, spike_fn)) for _ in in_channels)
            self.reg_convs = nn.ModuleList(nn.Sequential(FusedConvBNPLIF(hid, hid, 3, 1, spike_fn),
                                                         FusedConvBNPLIF(hid, hid, 3, 1, spike_fn)) for _ in in_channels)
            for name in ("cls_preds", "reg_preds", "obj_preds"):      # lone Conv2d -> SeqToANNContainer (utils_snn.py:25-27)
                setattr(self, name, nn.ModuleList(SeqToANNContainer(m) for m in getattr(self, name)))

    def _pred(self, name, k) -> nn.Conv2d:
        m = getattr(self, name)[k]
        return m[0] if isinstance(m, SeqToANNContainer) else m

    def initialize_biases(self, prior_prob):                    # spiking_yolo_head.py:135-146
        for k in range(len(self.cls_preds)):
            for name in ("cls_preds", "obj_preds"):
                self._pred(name, k).bias.data.fill_(-math.log((1 - prior_prob) / prior_prob))

    @torch.no_grad()
    def run(self, feats):
        """feats: spikes ``[T, B, H, W, C]`` fp16 per level -> ``[B, n_anchors, 5 + num_classes]`` fp32."""
        if self.training:
            raise RuntimeError("SpikingYOLOXHead (fused) is the inference path; call .eval()")
        dev = feats[0].device
        if not self.full_spike:
            rates = []
            for f in feats:
                _, B, H, W, C = f.shape
                rates.append(time_mean_planes(f.contiguous(), _new_planes(B, H, W, C, dev)))
            return super().run(rates)
        T = feats[0].shape[0]
        towers = []
        for k, f in enumerate(feats):
            x = self.stems[k].run(f.contiguous(), T)
            cls_feat = self.cls_convs[k][1].run(self.cls_convs[k][0].run(x, T), T)
            reg_feat = self.reg_convs[k][1].run(self.reg_convs[k][0].run(x, T), T)
            _, B, H, W, C = cls_feat.shape
            towers.append((time_mean_planes(cls_feat, _new_planes(B, H, W, C, dev)),
                           time_mean_planes(reg_feat, _new_planes(B, H, W, C, dev))))
        return self._predict_and_decode(towers)

    def forward(self, xin, labels=None, imgs=None):
        """Reference-shaped entry: ``xin`` = spikes ``[T, B, C, H, W]`` per level."""
        return self.run([x.permute(0, 1, 3, 4, 2).to(ACT_DTYPE).contiguous() for x in xin])


def nchw_to_planes(x: torch.Tensor) -> torch.Tensor:
    from .fused import split_f16
    return split_f16(x.permute(0, 2, 3, 1).contiguous().unsqueeze(0), 2)


# ------------------------------------------------------------------------------------------------
# whole model
# ------------------------------------------------------------------------------------------------
class SpikingYOLOX(nn.Module):
    """``SpikingYOLOX`` (spiking_yolox.py:23-74), inference: ``forward(x)`` with ``x`` the micro-bin histograms
    ``[B, Tm, 2, H, W]`` (or whatever the embedding takes) returns ``[B, n_anchors, 5 + num_classes]``."""

    def __init__(self, backbone, head: YOLOXHead, embedding=None, T=4):
        super().__init__()
        self.nb_steps = T
        self.embedding = embedding
        self.backbone = backbone
        self.head = head
        if backbone.backbone.T != T:
            raise ValueError("backbone was built for T=%d" % backbone.backbone.T)

    def set_ann_precision(self, precision: str = "fp32"):
        """Arithmetic of the ANN pyramid / head (the spiking backbone is unaffected: its inputs are exact in fp16).
        ``"fp32"`` (default): activations and weights as fp16 hi + lo planes, three product terms = fp32-equivalent,
        the reference's default evaluation.  ``"fp16"``: activations rounded to fp16 on their way into every conv
        (one input plane, two product terms, fp32 accumulation) -- what the reference computes under ``--fp16``
        autocast (tools/eval_event.py, yolox/evaluators/event_evaluator.py:180-190); predictions then agree with the
        fp32 ones to ~1e-2 relative, the reduced-precision bar of the parity contract."""
        if precision not in ("fp32", "fp16"):
            raise ValueError("ann precision must be 'fp32' or 'fp16'")
        for m in list(self.backbone.modules()) + list(self.head.modules()):
            if isinstance(m, (AnnBaseConv, YOLOXHead)):
                m.fp16_inputs = precision == "fp16"
        return self

    def embed(self, x, sampler=None):
        """spiking_yolox.py:41-57 up to the broadcast (which the backbone does implicitly for Ts == 1).  ``sampler``
        replaces the call of the sampler module (the raw-event front doors pass ``m.forward_events(...)``); the
        ``exp.norm`` variant ``ModuleList([embedding, BatchNorm2d(2)])`` (event_yolox_base.py:188-192) goes the same way."""
        call = sampler if sampler is not None else (lambda m: m(x))
        if isinstance(self.embedding, nn.ModuleList):
            x = call(self.embedding[0])
            if x.dim() > 4:
                x = x[0]
            if len(self.embedding) > 1:
                x = self.embedding[1](x)
        elif self.embedding is not None:
            x = call(self.embedding)
            if x.dim() > 5:
                x = x[0]
        if x.dim() == 4:
            x = x.unsqueeze(0)
        if x.shape[0] != 1 and x.shape[0] != self.nb_steps:
            raise AssertionError("the timestep of SNN is not matched with that of input")
        return x

    @torch.no_grad()
    def detect_frames(self, frames: torch.Tensor) -> torch.Tensor:
        """frames ``[Ts or T, B, 2, H, W]`` (sampler output) -> decoded predictions."""
        return self.head.run(self.backbone.run(frames))

    def forward(self, x, targets=None):
        if self.training:
            raise RuntimeError("SpikingYOLOX (fused) is the inference path; call .eval()")
        with torch.no_grad():
            return self.detect_frames(self.embed(x))

    @torch.no_grad()
    def forward_events(self, x, y, t, p, offsets, H: int, W: int, pad_to=None):
        """raw events -> detections: binning + sampler (``AdaptiveRSNNEmbedding.forward_events``), zero padding of
        the frames to ``pad_to = (H', W')`` (multiples of 32), detector."""
        frames = self.embed(None, sampler=lambda m: m.forward_events(x, y, t, p, offsets, H, W))
        return self.detect_frames(pad_frames(frames, pad_to))


def pad_frames(frames: torch.Tensor, size=None) -> torch.Tensor:
    """Zero-pad ``[..., H, W]`` at the bottom / right to ``size`` (default: next multiples of 32) -- the top-left
    letterbox placement of the reference's loaders (gen1.py:433-521) without the resize."""
    H, W = frames.shape[-2:]
    Hp, Wp = size if size is not None else ((H + 31) // 32 * 32, (W + 31) // 32 * 32)
    if (Hp, Wp) == (H, W):
        return frames
    return torch.nn.functional.pad(frames, (0, Wp - W, 0, Hp - H))


def build_syolox(depth, width, num_classes=2, T=3, embedding=None, spike_fn=None, in_channels=(256, 512, 1024),
                 use_spike=True):
    """The model of ``EventExp.get_model`` (event_yolox_base.py:188-218); e-yolox-s = (0.33, 0.50), e-yolox-m =
    (0.67, 0.75) (exps/default/e_yolox_s.py:13-14, e_yolox_m.py:13-14).  ``use_spike``: ``True`` (spiking backbone, ANN
    pyramid + head, :197-206), ``"full_spike"`` (spiking backbone + pyramid, ANN head on firing rates, :207-211) or
    ``"full_spike_v2"`` (spiking head towers too) -- the README trains and evaluates SYOLOX-M with ``full_spike``
    (readme.md:136-160)."""
    if use_spike is True or use_spike == "True":
        backbone = SpikingYOLOPAFPN(depth, width, in_channels=in_channels, in_dim=2, spike_fn=spike_fn, T=T)
        head = YOLOXHead(num_classes, width, in_channels=in_channels)
    elif isinstance(use_spike, str) and "full_spike" in use_spike:
        backbone = FullSpikeYOLOPAFPN(depth, width, in_channels=in_channels, in_dim=2, spike_fn=spike_fn, T=T)
        head = SpikingYOLOXHead(num_classes, width, in_channels=in_channels, spike_fn=spike_fn,
                                full_spike="v2" in use_spike, T=T)
    else:
        raise NotImplementedError("use_spike=%r: the fused detector covers True, 'full_spike', 'full_spike_v2'" % (use_spike,))
    model = SpikingYOLOX(backbone, head, embedding, T=T)
    head.initialize_biases(1e-2)
    return model


def postprocess(prediction, num_classes, conf_thre=0.7, nms_thre=0.45, class_agnostic=False):
    """``yolox/utils/boxes.py:33-77``: (cx, cy, w, h) -> corners, confidence filter, (class-wise) NMS.  Returns a
    list with one ``[n, 7]`` tensor (x1, y1, x2, y2, obj_conf, class_conf, class_pred) or None per image.
    Unlike the reference it does not modify ``prediction`` in place."""
    import torchvision
    pred = prediction.clone()
    pred[:, :, 0] = prediction[:, :, 0] - prediction[:, :, 2] / 2
    pred[:, :, 1] = prediction[:, :, 1] - prediction[:, :, 3] / 2
    pred[:, :, 2] = prediction[:, :, 0] + prediction[:, :, 2] / 2
    pred[:, :, 3] = prediction[:, :, 1] + prediction[:, :, 3] / 2
    output = [None] * len(pred)
    for i, image_pred in enumerate(pred):
        if not image_pred.size(0):
            continue
        class_conf, class_pred = torch.max(image_pred[:, 5:5 + num_classes], 1, keepdim=True)
        conf_mask = (image_pred[:, 4] * class_conf.squeeze(1) >= conf_thre)
        det = torch.cat((image_pred[:, :5], class_conf, class_pred.float()), 1)[conf_mask]
        if not det.size(0):
            continue
        if class_agnostic:
            keep = torchvision.ops.nms(det[:, :4], det[:, 4] * det[:, 5], nms_thre)
        else:
            keep = torchvision.ops.batched_nms(det[:, :4], det[:, 4] * det[:, 5], det[:, 6], nms_thre)
        output[i] = det[keep]
    return output
