"""Batch-1 latency of bench.py's `latency` block (SYOLOX-S, one 50 ms window -> predictions; detector as a CUDA graph) and the
B = 64 SYOLOX-M full_spike detector forward, for A/B runs through EAS_B200_LIB."""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import eas_snn_b200 as eas
from eas_snn_b200 import detector, fused
dev = torch.device("cuda:0")
H, W, TM = bench.H, bench.W, bench.TM
torch.manual_seed(80)
model = eas.AdaptiveRSNNEmbedding(**bench.SAMPLER_KW).to(dev).eval()
b = [torch.from_numpy(a).to(dev) for a in bench.host_batches(0, bench.BATCH)[0]]
def build(dw, mode):
    torch.manual_seed(83)
    d = detector.build_syolox(dw[0], dw[1], num_classes=2, T=3, embedding=model, use_spike=mode).to(dev).eval()
    for m in d.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.bias.data.fill_(0.6)
    return d
pad = lambda fr: torch.nn.functional.pad(fr, (0, 320 - W, 0, 256 - H))
det_s = build((0.33, 0.50), True)
o1 = b[4][:2]; n1 = int(o1[-1]); ev1 = tuple(a[:n1] for a in b[:4])
with torch.no_grad():
    fr1 = pad(model(eas.bin_events(*ev1, o1, H, W, TM, dtype=torch.float32))).contiguous()
graph = fused.GraphedForward(det_s.detect_frames, fr1)
def lat():
    with torch.no_grad():
        fr = model(eas.bin_events(*ev1, o1, H, W, TM, dtype=torch.float32))
    return graph(pad(fr))
for _ in range(5): out = lat()
torch.cuda.synchronize()
ts, tg = [], []
for _ in range(40):
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); out = lat(); c.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(c))
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); graph.graph.replay(); c.record(); torch.cuda.synchronize(); tg.append(a.elapsed_time(c))
print("batch-1 latency %.4f ms (detector graph alone %.4f ms), checksum %.6f" % (float(np.median(ts)), float(np.median(tg)), float(out.float().abs().mean())))
det_m = build((0.67, 0.75), "full_spike")
with torch.no_grad():
    fr = pad(model(eas.bin_events(*b, H, W, TM, dtype=torch.uint8)))
    for _ in range(3): p = det_m.detect_frames(fr)
    torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): p = det_m.detect_frames(fr)
    c.record(); torch.cuda.synchronize()
print("SYOLOX-M full_spike detector, B = 64: %.4f ms per forward, checksum %.6f" % (a.elapsed_time(c) / 20, float(p.float().abs().mean())))
