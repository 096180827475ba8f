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
