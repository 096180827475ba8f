"""Scratch: time the fused spiking CSPDarknet (SYOLOX-S / M shapes) on the GPU box."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eas_snn_b200 as eas
from eas_snn_b200 import fused
dev = torch.device("cuda:0")
FLOP = {"S": 2.13e9, "M": 6.61e9}
for name, (dep, wid) in {"S": (0.33, 0.5), "M": (0.67, 0.75)}.items():
    torch.manual_seed(0)
    net = fused.SpikingCSPDarknet(dep, wid, in_dim=2, T=3).to(dev).eval()
    # make it fire: crude BN shift so every layer has activity
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.bias.data.fill_(0.6)
    for B in (8, 64):
        x = torch.rand(1, B, 2, 256, 320, device=dev) * 2
        for _ in range(3):
            outs = net(x)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); outs = net(x); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        rates = {k: round(float(v.float().mean()), 3) for k, v in outs.items()}
        print(name, "B", B, "ms %.3f" % ms, "frames/s %.0f" % (B / ms * 1e3), "TFLOP/s(1x) %.1f" % (FLOP[name] * 3 * B / ms / 1e9), rates)
