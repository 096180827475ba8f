"""Sampler forward timing on the bench workload: dtype x algo (CUDA events, 10 reps)."""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import eas_snn_b200 as eas
dev = torch.device("cuda:0")
torch.manual_seed(80)
model = eas.AdaptiveRSNNEmbedding(**bench.SAMPLER_KW).to(dev).eval()
b = [torch.from_numpy(a).to(dev) for a in bench.host_batches(0, bench.BATCH)[0]]
for dt in (torch.float32, torch.int32):
    hist = eas.bin_events(*b, bench.H, bench.W, bench.TM, dtype=dt)
    for algo in ("auto", "tensor", "tensor_split", "fp32"):
        model.algo = algo
        ts = []
        with torch.no_grad():
            for r in range(12):
                a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); out = model(hist); c.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(c))
        print(dt, algo, "ms/call min %.3f median %.3f max %.3f" % (min(ts[2:]), float(np.median(ts[2:])), max(ts[2:])), float(out.abs().sum()))
