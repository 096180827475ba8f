"""Repeat the sampler forward on identical inputs and compare bit for bit (race / stale-state probe)."""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eas_snn_b200 as eas
from eas_snn_b200 import synth
dev = torch.device("cuda:0")
for (B, H, W, lo, hi) in [(3, 64, 80, 2e4, 6e4), (64, 240, 304, 2e4, 2e5), (5, 37 * 2, 116, 1e4, 3e4)]:
    x, y, t, p, off = synth.make_batch(9, B, H, W, lo, hi)
    torch.manual_seed(2)
    m = eas.AdaptiveRSNNEmbedding(kernel_size=5, depth=2, nb_steps=4, thresh=1, vreset=0, Ts=1, write_zero=True,
                                  spike_attach=True).to(dev)
    d = [torch.from_numpy(a).to(dev) for a in (x, y, t, p, off)]
    rec = torch.from_numpy(eas.pack_records(x, y, t, p)).to(dev)
    ranges = torch.tensor([[off[b], off[b + 1]] for b in range(B)], dtype=torch.int64, device=dev)
    for algo in ("auto", "fp32"):
        m.algo = algo
        with torch.no_grad():
            ref = m.forward_events(*d, H, W).clone()
            bad = 0
            for it in range(100):
                junk = torch.randn(1 << 22, device=dev)          # churn the allocator / L2 between calls
                a = m.forward_dat(rec, ranges, H, W) if it % 2 else m.forward_events(*d, H, W)
                if not torch.equal(a, ref):
                    bad += 1
                    if bad <= 2:
                        diff = (a != ref)
                        print("  mismatch it=%d: %d elements, max |d| %.3e" % (it, int(diff.sum()), float((a - ref).abs().max())))
                del junk
        print("B=%d %dx%d algo=%s: %d / 100 repeats differ" % (B, H, W, algo, bad), flush=True)
