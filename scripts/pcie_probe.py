"""H2D / D2H bandwidth of pinned host memory on this box (explains the e2e bound in bench.py)."""
import json, torch
dev = torch.device("cuda:0")
out = {}
for mb in (4, 64, 256):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for name, (dst, src) in (("h2d", (d, h)), ("d2h", (h, d))):
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            dst.copy_(src, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        out["%s_%dMB_GBs" % (name, mb)] = round(n * 10 / a.elapsed_time(b) / 1e6, 2)
# both directions at once
n = 256 << 20
h1, h2 = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d1, d2 = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
out["duplex_each_dir_GBs"] = round(n * 10 / dt / 1e9, 2)
print(json.dumps(out))
