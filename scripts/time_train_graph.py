"""SYOLOX-S training step of bench.py (8 windows): eager vs CUDA-graph replay (fused.GraphedTrainStep); checks that
both take the parameters to the same place."""
import copy, os, sys, time, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import eas_snn_b200 as eas
from eas_snn_b200 import fused, parallel
dev = torch.device("cuda:0")
H, W, TM, TB = bench.H, bench.W, bench.TM, 8
b = [torch.from_numpy(a).to(dev) for a in bench.host_batches(0, bench.BATCH)[0]]
off = b[4][:TB + 1]; n = int(off[-1])
hist = eas.bin_events(b[0][:n], b[1][:n], b[2][:n], b[3][:n], off, H, W, TM, dtype=torch.float32)

def build():
    torch.manual_seed(82)
    emb = eas.AdaptiveRSNNEmbedding(**bench.SAMPLER_KW).to(dev).train()
    bb = fused.SpikingCSPDarknet(0.33, 0.50, in_dim=2, T=3).to(dev).train()
    for m in bb.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.bias.data.fill_(0.6)
    if "cl" in sys.argv:
        bb = bb.to(memory_format=torch.channels_last)
    params = list(emb.parameters()) + list(bb.parameters())
    return emb, bb, params

def loss_of(emb, bb):
    def f(h):
        fr = torch.nn.functional.pad(emb(h), (0, 320 - W, 0, 256 - H))
        outs = bb(fr)
        return sum((v.mean() - 0.2) ** 2 for v in outs.values())
    return f

# eager
emb, bb, params = build()
opt = torch.optim.Adam(params, lr=1e-4)
f = loss_of(emb, bb)
def eager():
    opt.zero_grad(set_to_none=True)
    loss = f(hist); loss.backward(); opt.step(); eas.reset_net(bb)
    return loss
for _ in range(3): eager()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): le = eager()
torch.cuda.synchronize(); t_eager = (time.perf_counter() - t0) / 10 * 1e3
# graphed, from the same initial state
emb2, bb2, params2 = build()
opt2 = torch.optim.Adam(params2, lr=1e-4, capturable=True, fused="fused" in sys.argv)
step = fused.GraphedTrainStep(loss_of(emb2, bb2), [hist], params2, opt2, after=lambda: eas.reset_net(bb2), warmup=3)
for _ in range(10): lg = step(hist)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): lg = step(hist)
torch.cuda.synchronize(); t_graph = (time.perf_counter() - t0) / 20 * 1e3
print("eager %.3f ms/step, graphed %.3f ms/step; loss eager %.6f graphed %.6f" % (t_eager, t_graph, float(le), float(lg)))
# same trajectory? 13 eager steps == warmup 3 + capture... (capture does not execute) + 10 -> compare after equal counts
emb3, bb3, params3 = build()
opt3 = torch.optim.Adam(params3, lr=1e-4)
f3 = loss_of(emb3, bb3)
for _ in range(3 + 30):
    opt3.zero_grad(set_to_none=True); l3 = f3(hist); l3.backward(); opt3.step(); eas.reset_net(bb3)
d = max(float((p.detach() - q.detach()).abs().max()) for p, q in zip(params2, params3))
print("after 33 steps: max |param graphed - eager| = %.3e, loss eager %.6f graphed %.6f" % (d, float(l3), float(lg)))
