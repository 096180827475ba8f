"""Eager vs CUDA-graph replay of the fused SYOLOX-M backbone at several batch sizes."""
import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eas_snn_b200 import fused
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = fused.SpikingCSPDarknet(0.67, 0.75, in_dim=2, T=3).to(dev).eval()
for m in net.modules():
    if isinstance(m, torch.nn.BatchNorm2d): m.bias.data.fill_(0.6)
def t(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(8):
        a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
for B in [int(v) for v in sys.argv[1:]] or [1, 8, 64]:
    x = torch.rand(1, B, 2, 256, 320, device=dev) * 2
    te = t(lambda: net(x)); print("M B=%d eager %.3f ms" % (B, te), flush=True)
    g = fused.GraphedForward(net, x); print("captured", flush=True)
    tg = t(lambda: g(x)); print("M B=%d graph %.3f ms" % (B, tg), flush=True)
