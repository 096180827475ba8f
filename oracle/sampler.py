"""Oracle (TEST INFRASTRUCTURE): the adaptive event sampler, dense PyTorch restatement.

Follows ``AdaptiveRSNNEmbedding`` (``yolox/models/embedding.py:79-226``; ``update``
``:132-139``, ``forward`` ``:141-226``) and the ``Rectangle`` surrogate
(``yolox/models/activation.py:17-30``).  The reference finds spikes with ``nonzero`` and
gathers/scatters through index lists; this restatement does the same arithmetic with
dense masks (no host syncs), which is what the CUDA kernel mirrors.  It is pinned
against the reference itself by ``tests/golden/make_golden.py`` (forward values and
parameter gradients).

Everything is differentiable through autograd, so the backward oracle is
``torch.autograd.grad`` of :func:`sampler_forward`.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


class RectangleFn(torch.autograd.Function):
    """activation.py:17-30 -- forward ``x > 0`` (strict), backward ``g * 1[|x| < 0.5/a] * a``, a = 1."""

    alpha = 1.0

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return x.gt(0).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * ((x.abs() < 0.5 / RectangleFn.alpha).to(g.dtype) * RectangleFn.alpha)


def conv_stack(x, weights, biases):
    """``build_conv`` (embedding.py:106-111): Conv(k, pad k//2) [+ ReLU + Conv]*."""
    n = len(weights)
    for i in range(n):
        k = weights[i].shape[-1]
        x = F.conv2d(x, weights[i], biases[i], padding=k // 2)
        if i + 1 < n:
            x = F.relu(x)
    return x


def sampler_forward(events: torch.Tensor,
                    in_w, in_b, gate_w, gate_b,
                    Ts: int = 1, thresh: float = 1.0, vreset: Optional[float] = 0.0,
                    readout: str = "sum", spike_attach: bool = False,
                    write_zero: bool = False, use_abs: bool = False,
                    return_state: bool = False):
    """Dense restatement of ``AdaptiveRSNNEmbedding.forward`` (embedding.py:141-226).

    events : ``[B, Tm, C, H, W]`` or ``[B, Tl, Tm, C, H, W]`` micro-bin tensor.
    in_w/in_b, gate_w/gate_b : lists of conv weights / biases of ``input_conv`` and
    ``gate_conv`` (``depth`` entries each).
    Returns ``[Ts, B*, C, H, W]``.
    """
    if events.dim() > 5:                                  # :147-151
        events = events.flatten(end_dim=-5)
    ev = events.transpose(0, 1).flip(0)                   # :153-156 newest micro-bin first
    Tm = ev.shape[0]
    zero = torch.zeros_like(ev[0])
    s = zero                                              # spike_last :159
    vm = zero                                             # vmem       :160
    acc = zero                                            # vmem_avg   :166
    seg = torch.zeros_like(ev[0], dtype=torch.long)       # seg_ind    :165
    tl = torch.zeros_like(ev[0], dtype=torch.long) - 1    # t_last     :167
    agg = [torch.zeros_like(ev[0]) for _ in range(Ts)]    # aggregation :164
    t_exec = Tm
    for t in range(Tm):
        g_rec, c_rec = conv_stack(s, gate_w, gate_b).chunk(2, dim=-3)        # :171-172
        g_in, c_in = conv_stack(ev[t], in_w, in_b).chunk(2, dim=-3)          # :173-174
        gate = torch.sigmoid(g_in + g_rec)                                   # :175
        cur = c_in + c_rec                                                   # :176
        v = gate * vm + cur                                                  # :133
        s = RectangleFn.apply(v - thresh)                                    # :134
        if vreset is None:
            vm = v - thresh * s                                              # :136
        else:
            vm = v * (1 - s) + vreset * s                                    # :138
        acc = acc + v                                                        # :179
        sb = s.detach() > 0
        valid = sb & (seg < Ts)                                              # :181-184
        if readout == "sum":
            val = acc                                                        # :186
        elif readout == "last":
            val = vm                                                         # :188
        elif readout == "avg":
            val = acc / (t - tl).clamp(min=1).to(acc.dtype)                  # :190-191 (only read where valid)
        else:
            raise NotImplementedError(readout)
        if spike_attach:
            val = val * s                                                    # :192-193 (SAT)
        for k in range(Ts):                                                  # :194
            agg[k] = agg[k] + torch.where(valid & (seg == k), val, torch.zeros_like(val))
        tl = torch.where(valid, torch.full_like(tl, t), tl)                  # :196
        seg = seg + valid.long()                                             # :195
        acc = torch.where(sb, torch.zeros_like(acc), acc)                    # :197
        if int(seg.min()) >= Ts:                                             # :200-201
            t_exec = t + 1
            break
    ns = (~(s.detach() > 0)) & (seg < Ts)                                    # :203-206
    if readout == "sum":
        val = acc
    elif readout == "last":
        val = vm
    else:
        den = torch.where(ns, Tm - 1 - tl, torch.ones_like(tl))              # :211-214 (only read where ns)
        val = acc / den.to(acc.dtype)
    if write_zero:
        val = val * 0                                                        # :215-216 (RPD)
    for k in range(Ts):                                                      # :217
        agg[k] = agg[k] + torch.where(ns & (seg == k), val, torch.zeros_like(val))
    out = torch.stack(agg, 0)
    if use_abs:
        out = F.relu(out)                                                    # :218-220
    if return_state:
        return out, {"seg": seg, "t_last": tl, "vm": vm, "acc": acc, "s": s, "t_exec": t_exec}
    return out


class OracleSampler(nn.Module):
    """Module form with the reference's parameter layout (state-dict keys
    ``gate_conv.{0,2}.*`` / ``input_conv.{0,2}.*``) and init (embedding.py:106-127)."""

    def __init__(self, kernel_size, in_channel=2, out_channel=2, Ts=1, split=False, spike_attach=False,
                 write_zero=False, abs=False, depth=1, readout="sum", **kwargs_spikes):
        super().__init__()
        self.Ts, self.abs, self.readout = Ts, abs, readout
        self.write_zero, self.spike_attach = write_zero, spike_attach
        self.nb_steps = kwargs_spikes["nb_steps"] if "Tm" not in kwargs_spikes else kwargs_spikes["Tm"]
        self.thresh = kwargs_spikes["thresh"]
        self.vreset = kwargs_spikes["vreset"]
        self.depth = int(depth)
        self.gate_conv = self._build(out_channel, out_channel * 2, kernel_size, self.depth)
        self.input_conv = self._build(in_channel, out_channel * 2, kernel_size, self.depth)
        for m in self.input_conv.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.orthogonal_(m.weight, gain=nn.init.calculate_gain("relu"))
        for m in self.gate_conv.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, nonlinearity="sigmoid")

    @staticmethod
    def _build(cin, cout, k, depth):
        convs = [nn.Conv2d(cin, cout, k, padding=k // 2)]
        for _ in range(depth - 1):
            convs += [nn.ReLU(inplace=True), nn.Conv2d(cout, cout, k, padding=k // 2)]
        return nn.Sequential(*convs)

    def _wb(self, seq):
        convs = [m for m in seq if isinstance(m, nn.Conv2d)]
        return [c.weight for c in convs], [c.bias for c in convs]

    def forward(self, events):
        if events.dim() < 5:                                                 # embedding.py:144-146
            events, _ = torch.broadcast_tensors(events, torch.zeros((self.Ts,) + events.shape))
            return events
        iw, ib = self._wb(self.input_conv)
        gw, gb = self._wb(self.gate_conv)
        return sampler_forward(events, iw, ib, gw, gb, Ts=self.Ts, thresh=self.thresh, vreset=self.vreset,
                               readout=self.readout, spike_attach=self.spike_attach,
                               write_zero=self.write_zero, use_abs=self.abs)


def spike_count(events: torch.Tensor, nb_steps: int) -> torch.Tensor:
    """``SpikeCountEmbedding.forward`` (``yolox/models/embedding.py:14-24``): micro-bin histograms summed over the
    micro-bin axis; a single frame (< 5-D) is broadcast ``nb_steps`` times first (:15-16).  Pinned by
    ``tests/golden/count.npz`` (the reference class run on 5-D, 6-D and 4-D inputs)."""
    if events.dim() < 5:
        events = events.unsqueeze(0).expand((nb_steps,) + tuple(events.shape))
    elif events.dim() > 5:
        events = events.flatten(end_dim=-5).transpose(0, 1)
    else:
        events = events.transpose(0, 1)
    return events.sum(dim=0)


def recurrent_layer_forward(events: torch.Tensor, in_w, in_b, gate_w, gate_b, thresh: float = 1.0,
                            vreset: Optional[float] = 0.0, readout: str = "sum", relu: bool = False):
    """Dense restatement of ``SpikingEmbedding.forward`` (yolox/models/embedding.py:285-316; ``update`` :276-283): the
    sampler's gated recurrent spiking layer read out as sum_t v_t (pre-reset) or the last membrane potential."""
    if events.dim() > 5:
        events = events.flatten(end_dim=-5)
    ev = events.transpose(0, 1).flip(0)                                      # :292-297 newest micro-bin first
    s = vm = torch.zeros_like(ev[0])
    vsum = 0
    for t in range(ev.shape[0]):
        g_rec, c_rec = conv_stack(s, gate_w, gate_b).chunk(2, dim=-3)        # :303-304
        g_in, c_in = conv_stack(ev[t], in_w, in_b).chunk(2, dim=-3)          # :298-299 (tdLayer: same conv per step)
        v = torch.sigmoid(g_in + g_rec) * vm + (c_in + c_rec)                # :305-307, :277
        s = RectangleFn.apply(v - thresh)                                    # :278
        vm = v - thresh * s if vreset is None else v * (1 - s) + vreset * s  # :279-282
        vsum = vsum + v                                                      # :308
    out = vsum if readout == "sum" else vm                                   # :313-316
    return F.relu(out) if relu else out


def lif_layer_forward(events: torch.Tensor, w, b, decay, thresh: float = 1.0, vreset: Optional[float] = 0.0,
                      readout: str = "sum"):
    """Dense restatement of ``LIFEmbedding.forward`` (embedding.py:53-76) over ``LIFCell.forward`` (cell.py:37-65):
    conv stack per step, ``v = sigmoid(decay) * v + psp``, Rectangle spike, soft / hard reset; sum_t v_t or last v."""
    if events.dim() > 5:
        events = events.flatten(end_dim=-5)
    ev = events.transpose(0, 1).flip(0)                                      # :60-63
    vm = torch.zeros_like(ev[0])
    vsum = 0
    for t in range(ev.shape[0]):
        v = torch.sigmoid(decay) * vm + conv_stack(ev[t], w, b)              # cell.py:48, embedding.py:66
        s = RectangleFn.apply(v - thresh)                                    # cell.py:55
        vm = v - thresh * s if vreset is None else v * (1 - s) + vreset * s  # cell.py:58-61
        vsum = vsum + v                                                      # embedding.py:71
    return vsum if readout == "sum" else vm


def sampler_records(events: torch.Tensor, in_w, in_b, gate_w, gate_b, Ts: int = 1, thresh: float = 1.0,
                    vreset: Optional[float] = 0.0):
    """The analysis outputs of the adaptive sampler (embedding.py:180, 198-201, 221-224): ``t_last`` after every
    executed step ``[steps, B, 2, H, W]`` and the concatenated sub-threshold potentials of every executed step."""
    if events.dim() > 5:
        events = events.flatten(end_dim=-5)
    ev = events.transpose(0, 1).flip(0)
    s = vm = torch.zeros_like(ev[0])
    seg = torch.zeros_like(ev[0], dtype=torch.long)
    tl = torch.zeros_like(ev[0], dtype=torch.long) - 1
    t_rec, v_rec = [], []
    for t in range(ev.shape[0]):
        g_rec, c_rec = conv_stack(s, gate_w, gate_b).chunk(2, dim=-3)
        g_in, c_in = conv_stack(ev[t], in_w, in_b).chunk(2, dim=-3)
        v = torch.sigmoid(g_in + g_rec) * vm + (c_in + c_rec)
        s = (v - thresh > 0).to(v.dtype)
        vm = v - thresh * s if vreset is None else v * (1 - s) + vreset * s
        v_rec.append(v[(1 - s).bool()])                                      # :180
        valid = (s > 0) & (seg < Ts)
        seg = seg + valid.long()
        tl = torch.where(valid, torch.full_like(tl, t), tl)
        t_rec.append(tl.clone())                                             # :198-199
        if int(seg.min()) >= Ts:                                             # :200-201
            break
    return torch.stack(t_rec, 0), torch.cat(v_rec)
