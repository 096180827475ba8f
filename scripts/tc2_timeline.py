"""Development aid: timeline of the row-folded sampler kernel's warp roles on CTA 0 (step t = 1).
Needs a library built with -DEAS_TC2_TRACE:
  EAS_NVCC_EXTRA=-DEAS_TC2_TRACE EAS_B200_OBJDIR=/tmp/eas_trace_obj EAS_B200_LIB_OUT=eas_snn_b200/lib/libeas_b200_trace.so python -m eas_snn_b200.build
  EAS_B200_LIB=eas_snn_b200/lib/libeas_b200_trace.so python scripts/tc2_timeline.py"""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import eas_snn_b200 as eas
from eas_snn_b200 import _lib

dev = torch.device("cuda:0")
torch.manual_seed(80)
model = eas.AdaptiveRSNNEmbedding(**bench.SAMPLER_KW).to(dev).eval()
model.algo = "tensor"
b = [torch.from_numpy(a).to(dev) for a in bench.host_batches(0, bench.BATCH)[0]]
hist = eas.bin_events(*b, bench.H, bench.W, bench.TM, dtype=torch.float32)
lib = _lib.lib()
rec = np.zeros(5 * 2048, dtype=np.uint64)
cnt = np.zeros(5, dtype=np.int32)
fn = lib.eas_debug_tc2_trace
fn.restype = ctypes.c_int
with torch.no_grad():
    for _ in range(3):
        out = model(hist)
        per = fn(rec.ctypes.data_as(ctypes.c_void_p), cnt.ctypes.data_as(ctypes.c_void_p))
names = ["PROD", "MMA1", "MMA2", "E1", "E2"]
evn = {0: ["start", "x0_empty ok", "filled"], 1: ["start", "x0_full ok", "d1_empty ok", "issued"],
       2: ["start", "x1_full ok", "d2_empty ok", "issued"], 3: ["start", "d1_full ok", "x1_empty ok", "stored"],
       4: ["tile start", "d2_full ok", "d2 released", "done"]}
t0 = None
rows = []
for r in range(5):
    for i in range(cnt[r]):
        v = int(rec[r * per + i])
        rows.append((v & 0xffffffffff, r, (v >> 56) & 0xff, (v >> 40) & 0xffff))
rows.sort()
t0 = rows[0][0]
print("counts", cnt.tolist(), "span", rows[-1][0] - t0, "cycles")
# per role: mean duration of each phase (event e-1 -> e) over tiles 4..
for r in range(5):
    ev = [x for x in rows if x[1] == r]
    by = {}
    for t, _, e, tile in ev:
        by.setdefault(tile, []).append((e, t))
    tiles = sorted(by)
    if not tiles:
        continue
    ph = {}
    last_end = None
    per_tile = []
    for tl in tiles:
        seq = sorted(by[tl], key=lambda x: x[1])
        for (e0, ta), (e1, tb) in zip(seq[:-1], seq[1:]):
            ph.setdefault((e0, e1), []).append(tb - ta)
        if last_end is not None:
            per_tile.append(seq[-1][1] - last_end)
        last_end = seq[-1][1]
    print(names[r], "tiles", len(tiles), "cycles/tile median", int(np.median(per_tile)) if per_tile else -1)
    for (e0, e1), v in sorted(ph.items()):
        print("    %-14s -> %-14s median %7d  mean %7d  max %7d" % (evn[r][e0], evn[r][e1], np.median(v), np.mean(v), max(v)))
if len(sys.argv) > 1:
    for t, r, e, tile in rows[: int(sys.argv[1])]:
        print("%9d %-5s tile %3d %s" % (t - t0, names[r], tile, evn[r][e]))
