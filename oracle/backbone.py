"""Oracle (TEST INFRASTRUCTURE): spiking CSPDarknet (conv -> BN -> PLIF), PyTorch restatement.

Follows the reference topology after ``convert_to_spiking`` (``yolox/utils/utils_snn.py:16-58``):
  * ``BaseConv`` = conv -> BN -> act (``yolox/models/network_blocks.py:31-56``) becomes
    ``SeqToANNContainer(Conv2d)`` -> multi-step BN -> ``ParametricLIFNode`` (utils_snn.py:25-53);
  * ``Focus`` (network_blocks.py:191-213) is wrapped whole and stays ANN (conv -> BN -> SiLU);
  * ``Bottleneck`` with SEW add (:81-104), ``SPPBottleneck`` (:125-147), ``CSPLayer`` (:150-188);
  * ``CSPDarknet`` (``yolox/models/darknet.py:97-180``), dep/wid multipliers 0.33/0.50 (S), 0.67/0.75 (M).
State-dict keys match the reference (``stem.0.conv.conv.weight``, ``dark2.0.conv.0.weight``,
``...bn.running_mean``, ``...act.w``) so reference weights load directly; pinned by a golden
vector produced from the reference model run through ``oracle/sj_shim``.

The neuron inside is ``oracle.plif`` (spikingjelly restatement, PARITY UNPINNED).
"""
from __future__ import annotations

import copy

import torch
import torch.nn as nn

from .plif import ATan, OraclePLIF


class _Seq(nn.Sequential):
    """[T, B, ...] -> flatten -> apply -> un-flatten."""

    def forward(self, x):
        y = super().forward(x.flatten(0, 1))
        return y.view([x.shape[0], x.shape[1]] + list(y.shape[1:]))


class _BNm(nn.BatchNorm2d):
    def forward(self, x):
        y = super().forward(x.flatten(0, 1))
        return y.view([x.shape[0], x.shape[1]] + list(y.shape[1:]))


def _plif(spike_fn):
    return OraclePLIF(init_tau=2.0, decay_input=False, v_threshold=1.0, v_reset=None,
                      surrogate_function=copy.deepcopy(spike_fn), detach_reset=False, step_mode="m")


class SpikingBaseConv(nn.Module):
    def __init__(self, cin, cout, k, stride, spike_fn):
        super().__init__()
        self.conv = _Seq(nn.Conv2d(cin, cout, k, stride, (k - 1) // 2, bias=False))
        self.bn = _BNm(cout, eps=1e-3, momentum=0.03)     # init_yolo, event_yolox_base.py:179-183
        self.act = _plif(spike_fn)

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


class _AnnBaseConv(nn.Module):
    def __init__(self, cin, cout, k, stride):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride, (k - 1) // 2, bias=False)
        self.bn = nn.BatchNorm2d(cout, eps=1e-3, momentum=0.03)
        self.act = nn.SiLU()

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


class _Focus(nn.Module):
    def __init__(self, cin, cout, k):
        super().__init__()
        self.conv = _AnnBaseConv(cin * 4, cout, k, 1)

    def forward(self, x):
        a, b = x[..., ::2, ::2], x[..., 1::2, ::2]
        c, d = x[..., ::2, 1::2], x[..., 1::2, 1::2]
        return self.conv(torch.cat((a, b, c, d), dim=1))


class SpikingBottleneck(nn.Module):
    def __init__(self, cin, cout, shortcut, expansion, spike_fn):
        super().__init__()
        hid = int(cout * expansion)
        self.conv1 = SpikingBaseConv(cin, hid, 1, 1, spike_fn)
        self.conv2 = SpikingBaseConv(hid, cout, 3, 1, spike_fn)
        self.use_add = shortcut and cin == cout

    def forward(self, x):
        y = self.conv2(self.conv1(x))
        return y + x if self.use_add else y


class SpikingSPP(nn.Module):
    def __init__(self, cin, cout, spike_fn, ks=(5, 9, 13)):
        super().__init__()
        hid = cin // 2
        self.conv1 = SpikingBaseConv(cin, hid, 1, 1, spike_fn)
        self.m = nn.ModuleList([_Seq(nn.MaxPool2d(k, 1, k // 2)) for k in ks])
        self.conv2 = SpikingBaseConv(hid * (len(ks) + 1), cout, 1, 1, spike_fn)

    def forward(self, x):
        x = self.conv1(x)
        return self.conv2(torch.cat([x] + [m(x) for m in self.m], dim=-3))


class SpikingCSPLayer(nn.Module):
    def __init__(self, cin, cout, n, shortcut, spike_fn):
        super().__init__()
        hid = int(cout * 0.5)
        self.conv1 = SpikingBaseConv(cin, hid, 1, 1, spike_fn)
        self.conv2 = SpikingBaseConv(cin, hid, 1, 1, spike_fn)
        self.conv3 = SpikingBaseConv(2 * hid, cout, 1, 1, spike_fn)
        self.m = nn.Sequential(*[SpikingBottleneck(hid, hid, shortcut, 1.0, spike_fn) for _ in range(n)])

    def forward(self, x):
        return self.conv3(torch.cat((self.m(self.conv1(x)), self.conv2(x)), dim=-3))


class SpikingCSPDarknet(nn.Module):
    def __init__(self, dep_mul, wid_mul, in_dim=2, spike_fn=None, out_features=("dark3", "dark4", "dark5")):
        super().__init__()
        spike_fn = spike_fn if spike_fn is not None else ATan(2.0)
        c = int(wid_mul * 64)
        d = max(round(dep_mul * 3), 1)
        self.out_features = out_features
        self.stem = _Seq(_Focus(in_dim, c, 3))
        self.dark2 = nn.Sequential(SpikingBaseConv(c, c * 2, 3, 2, spike_fn),
                                   SpikingCSPLayer(c * 2, c * 2, d, True, spike_fn))
        self.dark3 = nn.Sequential(SpikingBaseConv(c * 2, c * 4, 3, 2, spike_fn),
                                   SpikingCSPLayer(c * 4, c * 4, d * 3, True, spike_fn))
        self.dark4 = nn.Sequential(SpikingBaseConv(c * 4, c * 8, 3, 2, spike_fn),
                                   SpikingCSPLayer(c * 8, c * 8, d * 3, True, spike_fn))
        self.dark5 = nn.Sequential(SpikingBaseConv(c * 8, c * 16, 3, 2, spike_fn),
                                   SpikingSPP(c * 16, c * 16, spike_fn),
                                   SpikingCSPLayer(c * 16, c * 16, d, False, spike_fn))

    def forward(self, x, return_all=False):
        outs = {}
        x = self.stem(x)
        outs["stem"] = x
        for name in ("dark2", "dark3", "dark4", "dark5"):
            x = getattr(self, name)(x)
            outs[name] = x
        if return_all:
            return outs
        return {k: v for k, v in outs.items() if k in self.out_features}

    def reset(self):
        pass


def reset_net(net):
    for m in net.modules():
        if isinstance(m, OraclePLIF):
            m.reset()


def calibrate_bn(net: nn.Module, x: torch.Tensor, seed: int = 0):
    """Make the random-init backbone fire (hard part 8): set BN running stats
    from the fixture input (momentum 1.0, two train-mode passes), perturb gamma/beta, go to eval."""
    g = torch.Generator().manual_seed(seed)
    bns = [m for m in net.modules() if isinstance(m, nn.BatchNorm2d)]
    for m in bns:
        m.weight.data = torch.empty_like(m.weight).uniform_(0.8, 1.2, generator=g)
        m.bias.data = torch.empty_like(m.bias).normal_(0.0, 0.1, generator=g)
    old = [m.momentum for m in bns]
    for m in bns:
        m.momentum = 1.0
    net.train()
    with torch.no_grad():
        for _ in range(2):
            net(x)
            reset_net(net)
    for m, o in zip(bns, old):
        m.momentum = o
    net.eval()
    return net


def fold_bn(conv_w: torch.Tensor, bn: nn.BatchNorm2d):
    """``fuse_conv_and_bn`` formula (yolox/utils/model_utils.py:61-75): returns (w', b')."""
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    return conv_w * scale.view(-1, 1, 1, 1), bn.bias - bn.running_mean * scale
