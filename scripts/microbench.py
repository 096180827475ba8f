"""Scratch kernel timings on the GPU box (CUDA events, L2 flushed between iterations)."""
import json
import sys
import os
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eas_snn_b200 as eas
from eas_snn_b200 import synth

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


res = {}
PEAK = 6556.5
# ---- binning, Gen1 batch 64
H, W = synth.GEN1
arrs = synth.gen1_batch(64)
d = [torch.from_numpy(a).to(dev) for a in arrs]
n = int(arrs[4][-1])
for s in ("tiles", "reds"):
    out = torch.empty((64, 4, 2, H, W), dtype=torch.int32, device=dev)
    med, mn = timeit(lambda: eas.bin_events(*d, H, W, 4, strategy=s, out=out))
    byt = 5 * n + out.numel() * 4
    res["bin_gen1_b64_" + s] = dict(ms=med, ms_min=mn, events=n, Mev_s=n / med / 1e3, GBs_5B=byt / med / 1e6,
                                    frac=byt / med / 1e6 / PEAK)
# ---- binning sweep single window
for (HH, WW) in (synth.GEN1, synth.MPX):
    for N in (10**5, 10**6, 10**7, 10**8):
        rng = np.random.default_rng(1)
        x, y, t, p = synth.make_window(rng, N, HH, WW)
        dd = [torch.from_numpy(a).to(dev) for a in (x, y, t, p, np.array([0, N], np.int64))]
        out = torch.empty((1, 4, 2, HH, WW), dtype=torch.int32, device=dev)
        for s in (("tiles", "reds") if HH == 240 else ("reds",)):
            med, mn = timeit(lambda: eas.bin_events(*dd, HH, WW, 4, strategy=s, out=out), iters=5)
            byt = 5 * N + out.numel() * 4
            res["bin_%dx%d_N%.0e_%s" % (HH, WW, N, s)] = dict(ms=med, Mev_s=N / med / 1e3, GBs_5B=byt / med / 1e6)
        del dd
# ---- sampler fwd
torch.manual_seed(80)
m = eas.AdaptiveRSNNEmbedding(kernel_size=5, depth=2, nb_steps=4, thresh=1, vreset=0, Ts=1, write_zero=True,
                              spike_attach=True).to(dev)
for B in (1, 8, 64):
    hist = eas.bin_events(*[torch.from_numpy(a).to(dev) for a in synth.gen1_batch(B)], H, W, 4, dtype=torch.float32)
    with torch.no_grad():
        med, mn = timeit(lambda: m(hist))
    flop = 2400.0 * H * W * 4 * B * 1.0
    res["sampler_fwd_B%d" % B] = dict(ms=med, ms_min=mn, TFLOPs=flop / med / 1e9, frames_s=B / med * 1e3)
# ---- plif
node = eas.ParametricLIFNode(decay_input=False, v_reset=None, surrogate_function=eas.ATan(2.0), step_mode="m").to(dev)
node.keep_v = False
for dt in (torch.float32, torch.bfloat16):
    x = (torch.rand((3, 64, 96, 64, 80), device=dev) * 1.5).to(dt)
    with torch.no_grad():
        med, mn = timeit(lambda: node(x))
    byt = x.numel() * x.element_size() * 2
    res["plif_fwd_%s" % str(dt)[6:]] = dict(ms=med, GBs=byt / med / 1e6, frac=byt / med / 1e6 / PEAK)
    xg = x.clone().requires_grad_(True)
    s = node(xg)
    go = torch.rand_like(s)
    med, mn = timeit(lambda: torch.autograd.grad(s, xg, go, retain_graph=True))
    byt = x.numel() * x.element_size() * 3
    res["plif_bwd_%s" % str(dt)[6:]] = dict(ms=med, GBs=byt / med / 1e6, frac=byt / med / 1e6 / PEAK)
for k, v in res.items():
    print(k, json.dumps({a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()}))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/microbench.json", "w"), indent=1)
