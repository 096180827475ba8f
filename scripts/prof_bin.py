"""A few binning calls on the bench workload (for ncu captures): compact byte histogram, then dense fp32."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import eas_snn_b200 as eas
dev = torch.device("cuda:0")
sets = [[torch.from_numpy(a).to(dev) for a in b] for b in bench.host_batches(0, bench.BATCH)]
for dt in (torch.uint8, torch.float32):
    for r in range(3):
        h = eas.bin_events(*sets[r], bench.H, bench.W, bench.TM, dtype=dt)
torch.cuda.synchronize()
print("ok")
