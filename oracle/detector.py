"""Oracle (TEST INFRASTRUCTURE): the whole SYOLOX detector (``use_spike True`` and the two ``full_spike`` variants),
PyTorch fp32 restatement.

Follows, with the reference's child names (so its ``state_dict`` loads with ``strict=True``):
  * ``SpikingYOLOPAFPN.forward`` (``yolox/models/spiking_yolo_pafpn.py:89-120``): spiking CSPDarknet
    (``oracle.backbone``), ``.mean(axis=0)`` over the T steps (:98), ANN top-down / bottom-up pyramid of
    ``BaseConv`` = conv -> BN -> SiLU and ``CSPLayer`` (``yolox/models/network_blocks.py:31-56, 81-104, 150-188``);
  * ``YOLOXHead.forward`` inference branch + ``decode_outputs`` (``yolox/models/yolo_head.py:141-199, 232-250``);
  * ``SpikingYOLOX.forward`` (``yolox/models/spiking_yolox.py:38-74``);
  * ``postprocess`` (``yolox/utils/boxes.py:33-77``) lives in the product (torchvision NMS) and is compared on
    identical predictions.
  * ``use_spike full_spike / full_spike_v2`` (``yolox/exp/event_yolox_base.py:207-211``): the whole ``YOLOPAFPN``
    (``yolo_pafpn.py:16-116``) converted by ``convert_to_spiking`` (``utils_snn.py:16-58``) and ``SpikingYOLOXHead``
    (``spiking_yolo_head.py:18-230``): time mean in front of an ANN head, or spiking towers with the time mean after
    the 1x1 predictors (:159-178).  PINNED by ``tests/golden/detector_full_spike{,_v2}.npz``.
PINNED by ``tests/golden/detector.npz``: the reference's own ``EventExp.get_model()`` (tiny width) run through
``oracle/sj_shim`` (the neuron inside stays the unpinned spikingjelly restatement, see ``oracle/plif.py``).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .backbone import SpikingBaseConv, SpikingCSPDarknet, SpikingCSPLayer, _AnnBaseConv, _Seq, reset_net


class _Bottleneck(nn.Module):
    def __init__(self, cin, cout, shortcut, expansion=0.5):
        super().__init__()
        hid = int(cout * expansion)
        self.conv1 = _AnnBaseConv(cin, hid, 1, 1)
        self.conv2 = _AnnBaseConv(hid, cout, 3, 1)
        self.use_add = shortcut and cin == cout

    def forward(self, x):
        y = self.conv2(self.conv1(x))
        return y + x if self.use_add else y


class _CSPLayer(nn.Module):
    def __init__(self, cin, cout, n, shortcut):
        super().__init__()
        hid = int(cout * 0.5)
        self.conv1 = _AnnBaseConv(cin, hid, 1, 1)
        self.conv2 = _AnnBaseConv(cin, hid, 1, 1)
        self.conv3 = _AnnBaseConv(2 * hid, cout, 1, 1)
        self.m = nn.Sequential(*[_Bottleneck(hid, hid, shortcut, 1.0) for _ in range(n)])

    def forward(self, x):
        return self.conv3(torch.cat((self.m(self.conv1(x)), self.conv2(x)), dim=1))


class OracleSpikingYOLOPAFPN(nn.Module):
    def __init__(self, depth, width, in_features=("dark3", "dark4", "dark5"), in_channels=(256, 512, 1024), in_dim=2,
                 spike_fn=None):
        super().__init__()
        self.backbone = SpikingCSPDarknet(depth, width, in_dim=in_dim, spike_fn=spike_fn, out_features=in_features)
        self.in_features = in_features
        c0, c1, c2 = (int(c * width) for c in in_channels)
        n = round(3 * depth)
        self.upsample = nn.Upsample(scale_factor=2, mode="nearest")
        self.lateral_conv0 = _AnnBaseConv(c2, c1, 1, 1)
        self.C3_p4 = _CSPLayer(2 * c1, c1, n, False)
        self.reduce_conv1 = _AnnBaseConv(c1, c0, 1, 1)
        self.C3_p3 = _CSPLayer(2 * c0, c0, n, False)
        self.bu_conv2 = _AnnBaseConv(c0, c0, 3, 2)
        self.C3_n3 = _CSPLayer(2 * c0, c1, n, False)
        self.bu_conv1 = _AnnBaseConv(c1, c1, 3, 2)
        self.C3_n4 = _CSPLayer(2 * c1, c2, n, False)

    def forward(self, x_seq):
        outs = self.backbone(x_seq)
        x2, x1, x0 = (outs[f].mean(axis=0) for f in self.in_features)
        fpn_out0 = self.lateral_conv0(x0)
        f_out0 = self.C3_p4(torch.cat([self.upsample(fpn_out0), x1], 1))
        fpn_out1 = self.reduce_conv1(f_out0)
        pan_out2 = self.C3_p3(torch.cat([self.upsample(fpn_out1), x2], 1))
        pan_out1 = self.C3_n3(torch.cat([self.bu_conv2(pan_out2), fpn_out1], 1))
        pan_out0 = self.C3_n4(torch.cat([self.bu_conv1(pan_out1), fpn_out0], 1))
        return pan_out2, pan_out1, pan_out0


class OracleYOLOXHead(nn.Module):
    def __init__(self, num_classes, width=1.0, strides=(8, 16, 32), in_channels=(256, 512, 1024)):
        super().__init__()
        self.num_classes, self.strides = num_classes, list(strides)
        hid = int(256 * width)
        self.stems = nn.ModuleList(_AnnBaseConv(int(c * width), hid, 1, 1) for c in in_channels)
        self.cls_convs = nn.ModuleList(nn.Sequential(_AnnBaseConv(hid, hid, 3, 1), _AnnBaseConv(hid, hid, 3, 1))
                                       for _ in in_channels)
        self.reg_convs = nn.ModuleList(nn.Sequential(_AnnBaseConv(hid, hid, 3, 1), _AnnBaseConv(hid, hid, 3, 1))
                                       for _ in in_channels)
        self.cls_preds = nn.ModuleList(nn.Conv2d(hid, num_classes, 1, 1, 0) for _ in in_channels)
        self.reg_preds = nn.ModuleList(nn.Conv2d(hid, 4, 1, 1, 0) for _ in in_channels)
        self.obj_preds = nn.ModuleList(nn.Conv2d(hid, 1, 1, 1, 0) for _ in in_channels)

    def forward(self, xin, decode=True):
        outs, grids, strides = [], [], []
        for k, x in enumerate(xin):
            x = self.stems[k](x)
            cls_feat, reg_feat = self.cls_convs[k](x), self.reg_convs[k](x)
            o = torch.cat([self.reg_preds[k](reg_feat), self.obj_preds[k](reg_feat).sigmoid(),
                           self.cls_preds[k](cls_feat).sigmoid()], 1)
            h, w = o.shape[-2:]
            yv, xv = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
            grids.append(torch.stack((xv, yv), 2).view(1, -1, 2).float())
            strides.append(torch.full((1, h * w, 1), float(self.strides[k])))
            outs.append(o.flatten(start_dim=2))
        out = torch.cat(outs, dim=2).permute(0, 2, 1)
        if not decode:
            return out
        g, s = torch.cat(grids, 1), torch.cat(strides, 1)
        return torch.cat([(out[..., 0:2] + g) * s, torch.exp(out[..., 2:4]) * s, out[..., 4:]], dim=-1)


class OracleFullSpikeYOLOPAFPN(nn.Module):
    """``convert_to_spiking(YOLOPAFPN(...))``: every BaseConv is conv -> BN -> PLIF on [T, B, C, H, W] (cat on dim -3)."""

    def __init__(self, depth, width, in_features=("dark3", "dark4", "dark5"), in_channels=(256, 512, 1024), in_dim=2,
                 spike_fn=None):
        super().__init__()
        self.backbone = SpikingCSPDarknet(depth, width, in_dim=in_dim, spike_fn=spike_fn, out_features=in_features)
        self.in_features = in_features
        c0, c1, c2 = (int(c * width) for c in in_channels)
        n = round(3 * depth)
        self.upsample = _Seq(nn.Upsample(scale_factor=2, mode="nearest"))
        self.lateral_conv0 = SpikingBaseConv(c2, c1, 1, 1, spike_fn)
        self.C3_p4 = SpikingCSPLayer(2 * c1, c1, n, False, spike_fn)
        self.reduce_conv1 = SpikingBaseConv(c1, c0, 1, 1, spike_fn)
        self.C3_p3 = SpikingCSPLayer(2 * c0, c0, n, False, spike_fn)
        self.bu_conv2 = SpikingBaseConv(c0, c0, 3, 2, spike_fn)
        self.C3_n3 = SpikingCSPLayer(2 * c0, c1, n, False, spike_fn)
        self.bu_conv1 = SpikingBaseConv(c1, c1, 3, 2, spike_fn)
        self.C3_n4 = SpikingCSPLayer(2 * c1, c2, n, False, spike_fn)

    def forward(self, x_seq):
        outs = self.backbone(x_seq)
        x2, x1, x0 = (outs[f] for f in self.in_features)
        fpn_out0 = self.lateral_conv0(x0)
        f_out0 = self.C3_p4(torch.cat([self.upsample(fpn_out0), x1], -3))
        fpn_out1 = self.reduce_conv1(f_out0)
        pan_out2 = self.C3_p3(torch.cat([self.upsample(fpn_out1), x2], -3))
        pan_out1 = self.C3_n3(torch.cat([self.bu_conv2(pan_out2), fpn_out1], -3))
        pan_out0 = self.C3_n4(torch.cat([self.bu_conv1(pan_out1), fpn_out0], -3))
        return pan_out2, pan_out1, pan_out0


class OracleSpikingYOLOXHead(OracleYOLOXHead):
    """``SpikingYOLOXHead``: ``full_spike=False`` averages every level over T in front of the ANN head
    (spiking_yolo_head.py:159-160); ``full_spike=True`` converts stems / towers to conv -> BN -> PLIF, wraps the 1x1
    predictors per time step and averages their outputs over T (:125-127, :173-178)."""

    def __init__(self, num_classes, width=1.0, strides=(8, 16, 32), in_channels=(256, 512, 1024), spike_fn=None,
                 full_spike=False):
        super().__init__(num_classes, width, strides, in_channels)
        self.full_spike = full_spike
        if full_spike:
            hid = int(256 * width)
            self.stems = nn.ModuleList(SpikingBaseConv(int(c * width), hid, 1, 1, spike_fn) for c in in_channels)
            self.cls_convs = nn.ModuleList(nn.Sequential(SpikingBaseConv(hid, hid, 3, 1, spike_fn),
                                                         SpikingBaseConv(hid, hid, 3, 1, spike_fn)) for _ in in_channels)
            self.reg_convs = nn.ModuleList(nn.Sequential(SpikingBaseConv(hid, hid, 3, 1, spike_fn),
                                                         SpikingBaseConv(hid, hid, 3, 1, spike_fn)) for _ in in_channels)
            for name in ("cls_preds", "reg_preds", "obj_preds"):
                setattr(self, name, nn.ModuleList(_Seq(m) for m in getattr(self, name)))

    def forward(self, xin, decode=True):
        if not self.full_spike:
            return super().forward([x.mean(axis=0) for x in xin], decode)
        outs, grids, strides = [], [], []
        for k, x in enumerate(xin):
            x = self.stems[k](x)
            cls_feat, reg_feat = self.cls_convs[k](x), self.reg_convs[k](x)
            o = torch.cat([self.reg_preds[k](reg_feat).mean(axis=0), self.obj_preds[k](reg_feat).mean(axis=0).sigmoid(),
                           self.cls_preds[k](cls_feat).mean(axis=0).sigmoid()], 1)
            h, w = o.shape[-2:]
            yv, xv = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
            grids.append(torch.stack((xv, yv), 2).view(1, -1, 2).float())
            strides.append(torch.full((1, h * w, 1), float(self.strides[k])))
            outs.append(o.flatten(start_dim=2))
        out = torch.cat(outs, dim=2).permute(0, 2, 1)
        if not decode:
            return out
        g, s = torch.cat(grids, 1), torch.cat(strides, 1)
        return torch.cat([(out[..., 0:2] + g) * s, torch.exp(out[..., 2:4]) * s, out[..., 4:]], dim=-1)


class OracleSpikingYOLOX(nn.Module):
    def __init__(self, depth, width, num_classes, T, embedding=None, spike_fn=None, use_spike=True):
        super().__init__()
        self.nb_steps = T
        self.embedding = embedding
        if use_spike is True:
            self.backbone = OracleSpikingYOLOPAFPN(depth, width, spike_fn=spike_fn)
            self.head = OracleYOLOXHead(num_classes, width)
        else:
            self.backbone = OracleFullSpikeYOLOPAFPN(depth, width, spike_fn=spike_fn)
            self.head = OracleSpikingYOLOXHead(num_classes, width, spike_fn=spike_fn, full_spike="v2" in use_spike)
        for ml in (self.head.cls_preds, self.head.obj_preds):
            for m in ml.modules():
                if isinstance(m, nn.Conv2d):
                    m.bias.data.fill_(-math.log((1 - 1e-2) / 1e-2))

    def detect_frames(self, frames):
        """frames [Ts or T, B, 2, H, W] -> decoded predictions [B, A, 5 + nc]."""
        if frames.shape[0] == 1:
            frames = frames.expand(self.nb_steps, -1, -1, -1, -1)
        assert frames.shape[0] == self.nb_steps
        out = self.head(self.backbone(frames.contiguous()))
        reset_net(self)
        return out

    def forward(self, x):
        f = self.embedding(x) if self.embedding is not None else x
        if f.dim() > 5:
            f = f[0]
        if f.dim() == 4:
            f = f.unsqueeze(0)
        return self.detect_frames(f)
