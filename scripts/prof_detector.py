"""SYOLOX-M whole-detector forward (B=64, T=3, 256x320) for ncu launch lists / timing.
usage: prof_detector.py [time]"""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eas_snn_b200 import detector
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = detector.build_syolox(0.67, 0.75, 2, 3).to(dev).eval()
for m in net.modules():
    if isinstance(m, torch.nn.BatchNorm2d):
        m.bias.data.fill_(0.6)
if os.environ.get("ANN") == "fp16":
    net.set_ann_precision("fp16")
B = int(os.environ.get("B", 64))
x = torch.rand(1, B, 2, 256, 320, device=dev) * 2
for _ in range(2):
    out = net.detect_frames(x)
torch.cuda.synchronize()
if len(sys.argv) > 1 and sys.argv[1] == "time":
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = net.detect_frames(x); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    bb = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); net.backbone.backbone(x); b.record(); torch.cuda.synchronize(); bb.append(a.elapsed_time(b))
    print("detector B=%d: %.3f ms (backbone %.3f ms) env=%s" % (B, float(np.median(ts)), float(np.median(bb)),
          {k: v for k, v in os.environ.items() if k.startswith("EAS_")}), flush=True)
