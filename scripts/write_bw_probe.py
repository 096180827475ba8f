"""What pure device writes cost on this box: memset / fill of the B = 64 Gen1 histogram (149 MB fp32, 37 MB u8), next to a
D2D copy -- the floor of a kernel that writes the histogram once."""
import torch
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, reps=20):
    ts = []
    for _ in range(reps):
        flush.zero_(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]
for mb in (37.4, 74.7, 149.4, 600.0):
    n = int(mb * 1e6)
    buf = torch.empty(n, dtype=torch.uint8, device=dev)
    src = torch.empty(n, dtype=torch.uint8, device=dev)
    ms = t(lambda: buf.zero_())
    ms2 = t(lambda: buf.copy_(src))
    print("%.1f MB: memset %.4f ms = %.0f GB/s written; copy %.4f ms = %.0f GB/s (read + write)" % (mb, ms, n / ms / 1e6, ms2, 2 * n / ms2 / 1e6))
