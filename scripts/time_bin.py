"""Back-to-back timing of the binning call on the bench workload: compact byte histogram and dense fp32."""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import eas_snn_b200 as eas
dev = torch.device("cuda:0")
sets = [[torch.from_numpy(a).to(dev) for a in b] for b in bench.host_batches(0, bench.BATCH)]
shape = (bench.BATCH, bench.TM, 2, bench.H, bench.W)
outs = {"u8": [eas.CompactHist.empty(shape, dev) for _ in range(4)],
        "f32": [torch.empty(shape, dtype=torch.float32, device=dev)]}
for name, bufs in outs.items():
    for r in range(4):
        eas.bin_events(*sets[r % 4], bench.H, bench.W, bench.TM, out=bufs[r % len(bufs)])
    ts = []
    for rep in range(7):
        torch.cuda.synchronize()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for r in range(40):
            eas.bin_events(*sets[r % 4], bench.H, bench.W, bench.TM, out=bufs[r % len(bufs)])
        c.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(c) / 40)
    print(os.environ.get("EAS_B200_LIB", "in-tree"), name, "bin ms/call min %.4f median %.4f" % (min(ts), float(np.median(ts))))
