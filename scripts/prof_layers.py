"""Stand-alone launches of the SYOLOX-M (B=64, T=3, 256x320) layers that dominate the forward, for ncu.
usage: prof_layers.py [time]"""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eas_snn_b200 import fused
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, T = int(os.environ.get("B", 64)), 3
g = torch.Generator(device=dev).manual_seed(1)

def spikes(*shape):
    return (torch.rand(shape, device=dev, generator=g) < 0.25).half()

def layer(cin, cout, k, s):
    m = fused.FusedConvBNPLIF(cin, cout, k, s).to(dev).eval()
    m.bn.bias.data.fill_(0.6)
    return m

cases = []
# name, module, input, kwargs
stem = fused._Focus(2, 48, 3).to(dev).eval()
frames = torch.rand(1, B, 2, 256, 320, device=dev) * 2
cases.append(("stem 8->48 k3 @128x160 (SiLU planes)", lambda: stem.run(frames)))
x_stem = stem.run(frames)
l0 = layer(48, 96, 3, 2)
cases.append(("dark2.0 48->96 k3 s2 (Tx=1, split input)", lambda: l0.run(x_stem, T, n_xsplit=2)))
x96 = spikes(T, B, 64, 80, 96)
l1 = layer(96, 48, 1, 1)
cases.append(("csp conv1 96->48 k1 @64x80", lambda: l1.run(x96, T)))
x48 = spikes(T, B, 64, 80, 48)
l2 = layer(48, 48, 3, 1)
cases.append(("bottleneck 48->48 k3 @64x80 (+res)", lambda: l2.run(x48, T, residual=x48)))
l3 = layer(96, 96, 1, 1)
cases.append(("csp conv3 96->96 k1 @64x80", lambda: l3.run(x96, T)))
l4 = layer(96, 192, 3, 2)
cases.append(("dark3.0 96->192 k3 s2 -> 32x40", lambda: l4.run(x96, T)))
x96b = spikes(T, B, 32, 40, 96)
l5 = layer(96, 96, 3, 1)
cases.append(("bottleneck 96->96 k3 @32x40 (+res)", lambda: l5.run(x96b, T, residual=x96b)))
x192 = spikes(T, B, 16, 20, 192)
l6 = layer(192, 192, 3, 1)
cases.append(("bottleneck 192->192 k3 @16x20 (+res)", lambda: l6.run(x192, T, residual=x192)))
for name, fn in cases:
    fn(); fn()
torch.cuda.synchronize()
if len(sys.argv) > 1 and sys.argv[1] == "time":
    for name, fn in cases:
        ts = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
        print("%-45s %8.1f us" % (name, float(np.median(ts))), flush=True)
