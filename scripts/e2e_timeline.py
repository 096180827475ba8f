"""Debug: per-stage device timestamps of bench.py's 3-stream e2e pipeline (H2D / compute / D2H)."""
import json, sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import eas_snn_b200 as eas
from eas_snn_b200.binning import HostEventBatch

dev = torch.device("cuda:0")
H, W, TM, TS, BATCH, NSETS = bench.H, bench.W, bench.TM, bench.TS, bench.BATCH, bench.NSETS
torch.manual_seed(80)
model = eas.AdaptiveRSNNEmbedding(**bench.SAMPLER_KW).to(dev).eval()
host = [HostEventBatch(*b) for b in bench.host_batches(0, BATCH)]
nmax = max(h.n for h in host)
s_in, s_cmp, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
slots = []
for _ in range(2):
    slots.append(dict(x=torch.empty(nmax, dtype=torch.int16, device=dev), y=torch.empty(nmax, dtype=torch.int16, device=dev),
                      t=torch.empty(nmax, dtype=torch.int64, device=dev), p=torch.empty(nmax, dtype=torch.uint8, device=dev),
                      off=torch.empty(BATCH + 1, dtype=torch.int64, device=dev),
                      hist=torch.empty((BATCH, TM, 2, H, W), dtype=torch.float32, device=dev),
                      host_out=torch.empty((TS, BATCH, 2, H, W), dtype=torch.float32).pin_memory(),
                      in_ready=torch.cuda.Event(), cmp_done=torch.cuda.Event(), out_done=torch.cuda.Event()))
ev = lambda: torch.cuda.Event(enable_timing=True)
def loop(steps, rec):
    for k in range(steps):
        sl, hb = slots[k % 2], host[k % NSETS]
        e = [ev() for _ in range(6)]
        with torch.cuda.stream(s_in):
            s_in.wait_event(sl["cmp_done"])
            e[0].record(s_in)
            sl["x"][:hb.n].copy_(hb.x, non_blocking=True); sl["y"][:hb.n].copy_(hb.y, non_blocking=True)
            sl["t"][:hb.n].copy_(hb.t, non_blocking=True); sl["p"][:hb.n].copy_(hb.p, non_blocking=True)
            sl["off"].copy_(hb.offsets, non_blocking=True)
            e[1].record(s_in); sl["in_ready"].record(s_in)
        with torch.cuda.stream(s_cmp):
            s_cmp.wait_event(sl["in_ready"])
            e[2].record(s_cmp)
            hist = eas.bin_events(sl["x"][:hb.n], sl["y"][:hb.n], sl["t"][:hb.n], sl["p"][:hb.n], sl["off"], H, W, TM, out=sl["hist"])
            with torch.no_grad():
                frames = model(hist)
            frames.record_stream(s_out)
            e[3].record(s_cmp); sl["cmp_done"].record(s_cmp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(sl["cmp_done"])
            e[4].record(s_out)
            sl["host_out"].copy_(frames, non_blocking=True)
            e[5].record(s_out); sl["out_done"].record(s_out)
        rec.append((e, time.perf_counter()))
loop(5, [])
torch.cuda.synchronize()
base = ev(); base.record(); t0 = time.perf_counter()
rec = []
loop(12, rec)
torch.cuda.synchronize()
print("wall ms/step", (time.perf_counter() - t0) * 1e3 / 12)
for trial in range(3):
    t1 = time.perf_counter(); loop(200, []); torch.cuda.synchronize()
    print("200 steps: wall ms/step", (time.perf_counter() - t1) * 1e3 / 200)
cs = bench.ClockSampler(0); cs.start(); time.sleep(0.3)
for trial in range(3):
    t1 = time.perf_counter(); loop(200, []); torch.cuda.synchronize()
    print("200 steps under nvidia-smi -lms 100: wall ms/step", (time.perf_counter() - t1) * 1e3 / 200)
print(cs.stop())
for k, (e, tc) in enumerate(rec):
    ts = [base.elapsed_time(x) for x in e]
    print("k=%2d host-issue %.2f | h2d %.2f-%.2f | cmp %.2f-%.2f | d2h %.2f-%.2f" % (k, (tc - t0) * 1e3, *ts))
