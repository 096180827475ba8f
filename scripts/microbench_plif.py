"""PLIF forward / backward kernels through the C ABI (no autograd overhead): HBM throughput against the
algorithmic bytes of SURVEY 8d (fwd 8 B, bwd 12 B per element-step in fp32; half in bf16)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eas_snn_b200 import _lib
dev = torch.device("cuda:0")
PEAK = 6556.5
L = _lib.lib()
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
T, N = 3, 64 * 96 * 64 * 80
w = torch.zeros((), device=dev)
for dt, code in ((torch.float32, _lib.EAS_F32), (torch.bfloat16, _lib.EAS_BF16)):
    x = (torch.rand((T, N), device=dev) * 1.5).to(dt)
    s, g, dx = torch.empty_like(x), torch.rand((T, N), device=dev).to(dt), torch.empty_like(x)
    gw = torch.zeros((), device=dev)
    cfg = _lib.PlifCfg(T=T, N=N, v_threshold=1.0, hard_reset=0, v_reset=0.0, decay_input=0, detach_reset=0,
                       surrogate=0, alpha=2.0, dtype=code)
    wsb = L.eas_plif_bwd_ws_bytes(C.byref(cfg))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    st = _lib.stream_ptr()
    ms = timeit(lambda: L.eas_plif_fwd(C.byref(cfg), _lib.ptr(x), _lib.ptr(w), None, _lib.ptr(s), None, st))
    byt = x.numel() * x.element_size() * 2
    print("plif_fwd kernel", dt, "ms %.4f GB/s %.0f frac %.3f" % (ms, byt / ms / 1e6, byt / ms / 1e6 / PEAK))
    ms = timeit(lambda: L.eas_plif_bwd(C.byref(cfg), _lib.ptr(x), _lib.ptr(w), None, _lib.ptr(g), _lib.ptr(dx),
                                       _lib.ptr(gw), _lib.ptr(ws), wsb, st))
    byt = x.numel() * x.element_size() * 3
    print("plif_bwd kernel", dt, "ms %.4f GB/s %.0f frac %.3f" % (ms, byt / ms / 1e6, byt / ms / 1e6 / PEAK))
