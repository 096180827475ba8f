"""Shared helpers for the parity tests (oracle = checker only)."""
from __future__ import annotations

import ast
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def sampler_case(z, name):
    cfg = ast.literal_eval(str(z[f"{name}/cfg"][0]))
    if cfg["vreset"] < -1e29:
        cfg["vreset"] = None
    params = {k.split("/param/")[1]: torch.from_numpy(z[k]) for k in z.files if k.startswith(f"{name}/param/")}
    grads = {k.split("/grad/")[1]: torch.from_numpy(z[k]) for k in z.files if k.startswith(f"{name}/grad/")}
    x = torch.from_numpy(z[f"{name}/x"].astype(np.float32))
    y = torch.from_numpy(z[f"{name}/y"])
    return cfg, params, grads, x, y


def sampler_kwargs(cfg):
    return dict(kernel_size=cfg["ksize"], in_channel=2, out_channel=2, readout=cfg["readout"], split=False,
                write_zero=cfg["write_zero"], abs=cfg["abs"], depth=cfg["depth"], nb_steps=cfg["Tm"],
                vreset=cfg["vreset"], thresh=1, embedding="arsnn", Ts=cfg["Ts"], spike_attach=cfg["spike_attach"])


def close_report(a: torch.Tensor, b: torch.Tensor, rtol=1e-5, atol=1e-5):
    """(ok, message).  Tolerance |a-b| <= atol + rtol*|b| (SURVEY.md section 7 hard part 3)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    nb = int(bad.sum())
    msg = "max|d|=%.3e  bad=%d/%d (%.2e)" % (float(err.max()) if err.numel() else 0.0, nb, err.numel(),
                                             nb / max(1, err.numel()))
    if nb:
        idx = torch.nonzero(bad)[:5].tolist()
        msg += "  first bad idx %s  got %s  want %s" % (idx, [float(a[tuple(i)]) for i in idx],
                                                        [float(b[tuple(i)]) for i in idx])
    return nb == 0, msg


def detector_case(z):
    """(meta, state_dict, hist) of tests/golden/detector.npz."""
    meta = ast.literal_eval(str(z["meta"][0]))
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    hist = torch.from_numpy(z["hist"].astype(np.float32))
    return meta, sd, hist


def detector_sampler_kwargs(meta):
    return dict(kernel_size=meta["ksize"], in_channel=2, out_channel=2, readout=meta["readout"], split=False,
                write_zero=meta["write_zero"], abs=meta["abs"], depth=meta["emb_depth"], nb_steps=meta["Tm"],
                vreset=meta["vreset"], thresh=meta["thresh"], embedding="arsnn", Ts=meta["Ts"],
                spike_attach=meta["spike_attach"])


def letterbox_cases(z):
    """(cfg tuple, regenerated input frames, golden strided sample, golden [sum, max, nonzero fraction]) per case."""
    i = 0
    while "%d/cfg" % i in z.files:
        ih, iw, h, w, center, lb = (int(v) for v in z["%d/cfg" % i])
        rng = np.random.default_rng(100 + i)
        fr = rng.poisson(0.7, (3, 2, ih, iw)).astype(np.float64)
        yield (ih, iw, h, w, bool(center), bool(lb)), fr, z["%d/sample" % i], z["%d/sum" % i]
        i += 1


# ------------------------------------------------------------------------------------------------
# A plain ANN CSPDarknet in the shape of the reference's classes (yolox/models/network_blocks.py:31-213,
# darknet.py:97-180: children .conv / .bn / .act, class name "Focus", concatenations on dim -3) -- the object a
# maintainer hands to ``convert_to_spiking`` (utils_snn.py:16-58).  Test stand-in: the reference package itself
# cannot travel to the GPU box.
# ------------------------------------------------------------------------------------------------
import torch.nn as _nn


class BaseConv(_nn.Module):
    def __init__(self, cin, cout, ksize, stride):
        super().__init__()
        self.conv = _nn.Conv2d(cin, cout, ksize, stride, (ksize - 1) // 2, bias=False)
        self.bn = _nn.BatchNorm2d(cout)
        self.act = _nn.SiLU(inplace=True)

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


class Focus(_nn.Module):
    def __init__(self, cin, cout, ksize=1):
        super().__init__()
        self.conv = BaseConv(cin * 4, cout, ksize, 1)

    def forward(self, x):
        tl, bl = x[..., ::2, ::2], x[..., 1::2, ::2]
        tr, br = x[..., ::2, 1::2], x[..., 1::2, 1::2]
        return self.conv(torch.cat((tl, bl, tr, br), dim=1))


class Bottleneck(_nn.Module):
    def __init__(self, cin, cout, shortcut=True, expansion=0.5):
        super().__init__()
        hid = int(cout * expansion)
        self.conv1 = BaseConv(cin, hid, 1, 1)
        self.conv2 = BaseConv(hid, cout, 3, 1)
        self.use_add = shortcut and cin == cout

    def forward(self, x):
        y = self.conv2(self.conv1(x))
        return y + x if self.use_add else y


class SPPBottleneck(_nn.Module):
    def __init__(self, cin, cout, ks=(5, 9, 13)):
        super().__init__()
        hid = cin // 2
        self.conv1 = BaseConv(cin, hid, 1, 1)
        self.m = _nn.ModuleList([_nn.MaxPool2d(k, 1, k // 2) for k in ks])
        self.conv2 = BaseConv(hid * (len(ks) + 1), cout, 1, 1)

    def forward(self, x):
        x = self.conv1(x)
        return self.conv2(torch.cat([x] + [m(x) for m in self.m], dim=-3))


class CSPLayer(_nn.Module):
    def __init__(self, cin, cout, n=1, shortcut=True):
        super().__init__()
        hid = int(cout * 0.5)
        self.conv1 = BaseConv(cin, hid, 1, 1)
        self.conv2 = BaseConv(cin, hid, 1, 1)
        self.conv3 = BaseConv(2 * hid, cout, 1, 1)
        self.m = _nn.Sequential(*[Bottleneck(hid, hid, shortcut, 1.0) for _ in range(n)])

    def forward(self, x):
        return self.conv3(torch.cat((self.m(self.conv1(x)), self.conv2(x)), dim=-3))


class AnnCSPDarknet(_nn.Module):
    def __init__(self, dep_mul, wid_mul, in_dim=2, out_features=("dark3", "dark4", "dark5")):
        super().__init__()
        c, d = int(wid_mul * 64), max(round(dep_mul * 3), 1)
        self.out_features = out_features
        self.stem = Focus(in_dim, c, 3)
        self.dark2 = _nn.Sequential(BaseConv(c, c * 2, 3, 2), CSPLayer(c * 2, c * 2, d, True))
        self.dark3 = _nn.Sequential(BaseConv(c * 2, c * 4, 3, 2), CSPLayer(c * 4, c * 4, d * 3, True))
        self.dark4 = _nn.Sequential(BaseConv(c * 4, c * 8, 3, 2), CSPLayer(c * 8, c * 8, d * 3, True))
        self.dark5 = _nn.Sequential(BaseConv(c * 8, c * 16, 3, 2), SPPBottleneck(c * 16, c * 16),
                                    CSPLayer(c * 16, c * 16, d, False))

    def forward(self, x):
        outs = {}
        x = self.stem(x)
        for name in ("dark2", "dark3", "dark4", "dark5"):
            x = getattr(self, name)(x)
            outs[name] = x
        return {k: v for k, v in outs.items() if k in self.out_features}
