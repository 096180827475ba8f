import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import eas_snn_b200 as eas
dev = torch.device("cuda:0")
torch.manual_seed(80)
model = eas.AdaptiveRSNNEmbedding(**bench.SAMPLER_KW).to(dev).eval()
sets = [[torch.from_numpy(a).to(dev) for a in b] for b in bench.host_batches(0, bench.BATCH)]
ref = []
with torch.no_grad():
    for s in sets:
        ref.append(model.forward_events(*s, bench.H, bench.W).clone())
    torch.cuda.synchronize()
    bad = 0
    for it in range(3000):
        out = model.forward_events(*sets[it % 4], bench.H, bench.W)
        if it % 97 == 0:
            bad += int(not torch.equal(out, ref[it % 4]))
    torch.cuda.synchronize()
    bad += int(not torch.equal(out, ref[(3000 - 1) % 4]))
eas.poll_compact(wait=True)
print("3000 back-to-back forward_events (bin + sampler, compact histogram, PDL launches): mismatches", bad)
