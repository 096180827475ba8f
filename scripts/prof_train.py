"""One eager SYOLOX-S training step of bench.py (8 windows, channels-last) inside a cudaProfiler range, for an ncu launch
list: ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ..."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import eas_snn_b200 as eas
from eas_snn_b200 import fused
dev = torch.device("cuda:0")
H, W, TM, TB = bench.H, bench.W, bench.TM, 8
b = [torch.from_numpy(a).to(dev) for a in bench.host_batches(0, bench.BATCH)[0]]
off = b[4][:TB + 1]; n = int(off[-1])
hist = eas.bin_events(b[0][:n], b[1][:n], b[2][:n], b[3][:n], off, H, W, TM, dtype=torch.float32)
torch.manual_seed(82)
emb = eas.AdaptiveRSNNEmbedding(**bench.SAMPLER_KW).to(dev).train()
bb = fused.SpikingCSPDarknet(0.33, 0.50, in_dim=2, T=3).to(dev).train()
if "nchw" not in sys.argv:
    bb = bb.to(memory_format=torch.channels_last)
for m in bb.modules():
    if isinstance(m, torch.nn.BatchNorm2d):
        m.bias.data.fill_(0.6)
params = list(emb.parameters()) + list(bb.parameters())
opt = torch.optim.Adam(params, lr=1e-4)
def step():
    opt.zero_grad(set_to_none=True)
    fr = torch.nn.functional.pad(emb(hist), (0, 320 - W, 0, 256 - H))
    loss = sum((v.mean() - 0.2) ** 2 for v in bb(fr).values())
    loss.backward(); opt.step(); eas.reset_net(bb)
for _ in range(3): step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
