"""Sampler forward timing on the bench workload: input form x algo; CUDA-event time of 16 back-to-back forwards."""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import eas_snn_b200 as eas
dev = torch.device("cuda:0")
torch.manual_seed(80)
model = eas.AdaptiveRSNNEmbedding(**bench.SAMPLER_KW).to(dev).eval()
b = [torch.from_numpy(a).to(dev) for a in bench.host_batches(0, bench.BATCH)[0]]
for dt in (torch.uint8, torch.float32, torch.int32):
    hist = eas.bin_events(*b, bench.H, bench.W, bench.TM, dtype=dt)
    for algo in ("auto", "tensor", "tensor_split", "fp32"):
        if dt == torch.uint8 and algo in ("tensor_split", "fp32"):
            continue
        model.algo = algo
        ts = []
        with torch.no_grad():
            for r in range(10):
                torch.cuda.synchronize()
                a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(16):
                    out = model(hist)
                c.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(c) / 16)
        print(dt, algo, "ms/forward min %.4f median %.4f" % (min(ts[2:]), float(np.median(ts[2:]))), float(out.abs().sum()))
