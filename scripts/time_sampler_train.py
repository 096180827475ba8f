"""Sampler training step (forward saving sequences + BPTT backward) on the bench workload, B = 8 and 64."""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import eas_snn_b200 as eas
from eas_snn_b200 import synth
dev = torch.device("cuda:0")
torch.manual_seed(80)
model = eas.AdaptiveRSNNEmbedding(**bench.SAMPLER_KW).to(dev).train()
for B in (8, 64):
    b = [torch.from_numpy(a).to(dev) for a in synth.gen1_batch(B)]
    hist = eas.bin_events(*b, bench.H, bench.W, bench.TM, dtype=torch.float32)
    def step():
        model.zero_grad(set_to_none=True)
        out = model(hist)
        out.sum().backward()
    for _ in range(3): step()
    torch.cuda.synchronize()
    tf, tb = [], []
    for _ in range(8):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        model.zero_grad(set_to_none=True)
        e[0].record(); out = model(hist); e[1].record(); out.sum().backward(); e[2].record()
        torch.cuda.synchronize()
        tf.append(e[0].elapsed_time(e[1])); tb.append(e[1].elapsed_time(e[2]))
    print("B=%d train fwd %.3f ms, bwd %.3f ms" % (B, float(np.median(tf)), float(np.median(tb))))
