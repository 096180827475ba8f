"""One sampler forward on the bench workload (for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import eas_snn_b200 as eas
dev = torch.device("cuda:0")
torch.manual_seed(80)
model = eas.AdaptiveRSNNEmbedding(**bench.SAMPLER_KW).to(dev).eval()
if len(sys.argv) > 1 and sys.argv[1] != "dense":
    model.algo = sys.argv[1]
b = [torch.from_numpy(a).to(dev) for a in bench.host_batches(0, bench.BATCH)[0]]
hist = eas.bin_events(*b, bench.H, bench.W, bench.TM, dtype=torch.float32 if "dense" in sys.argv else torch.uint8)
with torch.no_grad():
    for _ in range(3):
        out = model(hist)
torch.cuda.synchronize()
print(float(out.abs().sum()))
