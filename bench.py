#!/usr/bin/env python
"""bench.py -- headline benchmark of the EAS-SNN hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): "Mevents/s sampled+encoded" = events / (t_binning + t_sampler) on
BASELINE config[1]: Gen1 (240x304) windows at Gen1 event rate, batch 64 per GPU, Tm=4 micro-bins,
published sampler flags (depth 2, k 5, SAT + RPD, Ts=1).  One "step" = eas_bin_events +
eas_sampler_fwd over one batch of 64 synthetic windows.  Windows shard by sequence across ranks with
no collective -> weak scaling (64 windows per GPU).

  value : inputs already resident in HBM, CUDA-event timed, max over ranks
  e2e   : same step through the public module API from pinned HOST buffers: H2D of (x,y,t,p,offsets)
          and D2H of the sampled frames inside the timed region (3-stream pipeline)
  roofline / cpu_baseline / clocks: see DESIGN.md "Measurement"
--impl reference times the reference algorithm's CPU restatement (oracle/, numpy + torch CPU, all
host threads) on the same workload; the reference itself is Python that cannot travel to the GPU
box, so kind = "port".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, TM, TS, BATCH = 240, 304, 4, 1, 64
NSETS = 4  # rotating input batches: 4 x ~60 MB events (+150 MB histogram per step) > 126 MB L2
SAMPLER_KW = dict(kernel_size=5, in_channel=2, out_channel=2, readout="sum", split=False, write_zero=True,
                  abs=False, depth=2, nb_steps=TM, vreset=0, thresh=1, embedding="arsnn", Ts=TS, spike_attach=True)
WORKLOAD = "gen1_240x304_b64_Tm4_sampler_d2k5_sat_rpd"
METRIC, UNIT = "Mevents/s sampled+encoded", "Mevents/s"
# both arms print this config verbatim (the driver compares them); everything descriptive goes to "config_notes"
CONFIG = {"workload": WORKLOAD, "batch_per_gpu": BATCH, "H": H, "W": W, "Tm": TM, "Ts": TS}
MIN_TIMED_S = 1.2   # every timed region is repeated until it covers at least this long (clock samples: 100 ms)


_JSON_LINE: list = []     # the one line main() prints on the real stdout


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def host_batches(rank: int, batch: int):
    from eas_snn_b200 import synth
    return [synth.gen1_batch(batch, cfg=2, first_sample=(rank * NSETS + s) * batch) for s in range(NSETS)]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            # nvidia-smi takes a while to start (and holds driver locks while it does): wait for its
            # first sample so that start-up never overlaps a timed region
            t_end = time.time() + 10.0
            while time.time() < t_end and os.path.getsize(self.path) == 0:
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        for line in open(self.path):
            f = [c.strip() for c in line.split(",")]
            if len(f) >= 9:
                rows.append(f)
        os.unlink(self.path)
        if not rows:
            return out
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        out.update(sm_mhz=float(np.median(sm)) if sm else None,
                   sm_max_mhz=float(rows[0][2]) if rows[0][2].replace(".", "").isdigit() else None,
                   reasons=sorted(reasons), samples=len(rows))
        return out


# ------------------------------------------------------------------------------------------------
# CPU restatement of the reference path (oracle/): the cpu_baseline leg and the --impl reference arm
# ------------------------------------------------------------------------------------------------
def cpu_step(model, batch):
    from oracle import binning as ob
    x, y, t, p, off = batch
    hist = ob.micro_sum_batch(x, y, t, p, off, H, W, TM)           # gen1.py:313-360 (numpy, 1 thread)
    with torch.no_grad():
        frames = model(torch.from_numpy(hist).float())             # embedding.py:141-226 (torch CPU, all threads)
    return frames


def make_cpu_model():
    from oracle.sampler import OracleSampler
    torch.manual_seed(80)
    return OracleSampler(**SAMPLER_KW).eval()


def cpu_baseline(budget_s: float = 12.0, windows: int = 16):
    """Bounded sample: `windows` windows of the rank-0 workload, repeated until ~budget_s."""
    from eas_snn_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = make_cpu_model()
    batch = synth.gen1_batch(windows, cfg=2, first_sample=0)
    n = int(batch[4][-1])
    cpu_step(model, synth.gen1_batch(2, cfg=2, first_sample=0))   # warm-up
    times, t_end = [], time.perf_counter() + budget_s
    while len(times) < 2 or (time.perf_counter() < t_end and len(times) < 50):
        t0 = time.perf_counter()
        cpu_step(model, batch)
        times.append(time.perf_counter() - t0)
    med = float(np.median(times))
    return {"value": n / med / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d of the %d windows of one batch (%d events), median of %d passes; numpy binning "
                      "(1 thread) + torch-CPU sampler (%d threads)" % (windows, BATCH, n, len(times), cores)}


def cpu_frames_baseline(budget_s: float = 12.0, model: str = "m"):
    """The oracle port of the reference's detector forward on the host, batch 1, T=3, one 240x304 window zero-padded to
    256x320: events -> bins -> sampler -> detector -> decoded predictions.
      model "m": SYOLOX-M, use_spike full_spike -- BASELINE configs 2/3, the way the README evaluates it (readme.md:157-160)
      model "s": SYOLOX-S, use_spike True       -- BASELINE config 1 (the latency block's CPU counterpart)"""
    from eas_snn_b200 import synth
    from oracle import detector as odet
    from oracle.plif import ATan as OATan
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(80)
    dw, mode, name = ((0.67, 0.75), "full_spike", "SYOLOX-M (use_spike full_spike)") if model == "m" else \
        ((0.33, 0.50), True, "SYOLOX-S (use_spike True)")
    net = odet.OracleSpikingYOLOX(dw[0], dw[1], 2, 3, embedding=make_cpu_model(), spike_fn=OATan(2.0),
                                  use_spike=mode).eval()
    for mod in net.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.bias.data.fill_(0.6)
    batch = synth.gen1_batch(1, cfg=2, first_sample=0)      # the window the GPU latency block uses (rank 0, set 0)

    def one():
        with torch.no_grad():
            fr = cpu_step(net.embedding, batch)
            return net.detect_frames(torch.nn.functional.pad(fr, (0, 320 - W, 0, 256 - H)))

    one()
    times, t_end = [], time.perf_counter() + budget_s
    while len(times) < 2 or (time.perf_counter() < t_end and len(times) < 20):
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
    med = float(np.median(times))
    return {"value": 1.0 / med, "unit": "frames/s", "ms_per_frame": med * 1e3, "cores": cores, "kind": "port",
            "sample": "%s, batch 1, T=3, 256x320, %d events: numpy binning + torch-CPU sampler, spiking CSPDarknet, "
                      "pyramid, head, decode (%d threads), median of %d passes" % (name, int(batch[4][-1]), cores, len(times))}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from eas_snn_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = make_cpu_model()
    # bounded sample of the 64-window batch, sized so that K steps end within a couple of minutes
    windows = max(2, min(16, 480 // max(1, args.steps)))
    sets = [synth.gen1_batch(windows, cfg=2, first_sample=s * BATCH) for s in range(NSETS)]
    for w in range(args.warmup):
        cpu_step(model, sets[w % NSETS])
    n_ev, t0 = 0, time.perf_counter()
    for k in range(args.steps):
        cpu_step(model, sets[k % NSETS])
        n_ev += int(sets[k % NSETS][4][-1])
    dt = time.perf_counter() - t0
    val = n_ev / dt / 1e6
    sample = ("each step = %d of the %d windows of a batch; numpy binning (1 thread) + torch-CPU sampler "
              "(%d threads)" % (windows, BATCH, cores))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(CONFIG),
            "config_notes": {"note": "CPU restatement (oracle/) of the reference path; the Python reference cannot "
                                     "travel to the GPU box"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if not args.no_backbone:
        # the frames/s leg of the metric on the host: SYOLOX-M (full_spike), batch 1
        line["frames"] = cpu_frames_baseline(budget_s=10.0, model="m")
    _JSON_LINE.append(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import eas_snn_b200 as eas
    from eas_snn_b200 import parallel
    from eas_snn_b200.binning import HostEventBatch

    rank, world, local = parallel.init_distributed("nccl")
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    all_cpus = os.sched_getaffinity(0)
    numa = parallel.bind_to_gpu_numa_node(local)    # before any pinned host buffer is allocated
    peak_gbs, peak_src, sm_max_mhz = load_peaks()

    torch.manual_seed(80)
    model = eas.AdaptiveRSNNEmbedding(**SAMPLER_KW).to(dev).eval()
    host_np = host_batches(rank, BATCH)
    host = [HostEventBatch(*b) for b in host_np]
    devb = [hb.to_device(dev) for hb in host]
    hist_buf = torch.empty((BATCH, TM, 2, H, W), dtype=torch.float32, device=dev)
    # the histogram between the two kernels is the compact byte form (one byte per bin + the exact list of saturated
    # bins, include/eas_b200.h EAS_U8): the same information as the reference's dense tensor, a quarter of the bytes
    # to write and to read -- what forward_events / forward_dat use whenever the tensor-core sampler takes the call
    chist_buf = eas.CompactHist.empty((BATCH, TM, 2, H, W), dev)
    torch.cuda.synchronize()

    def step(db):
        hist = eas.bin_events(*db, H, W, TM, out=chist_buf)
        with torch.no_grad():
            return model(hist)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM ---------------------------------------------------------
    clocks = ClockSampler(local)
    clocks.start()                       # sampled through warm-up, value and e2e regions
    for w in range(max(args.warmup, 3)):
        step(devb[w % NSETS])
    barrier()
    def timed_value():
        """EXACTLY args.steps steps, bracketed by barrier + synchronize, CUDA events, max over ranks."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 0
        e0.record()
        for k in range(args.steps):
            step(devb[k % NSETS])
            n += host[k % NSETS].n
        e1.record()
        barrier()
        return parallel.max_over_ranks(e0.elapsed_time(e1), dev), n

    # the K-step region is repeated until the repeats cover MIN_TIMED_S (a 13 ms region says little and the clock
    # sampler never lands inside it); the reported figure is the MEDIAN repeat, `steps` stays what was asked for
    ms0, n_ev = timed_value()
    repeats = int(min(500, max(1, np.ceil(MIN_TIMED_S * 1e3 / max(ms0, 1e-3)))))
    if world > 1:
        repeats = int(parallel.max_over_ranks(float(repeats), dev))
    reps_ms = [ms0] + [timed_value()[0] for _ in range(repeats - 1)]
    ms = float(np.median(reps_ms))
    total_ev = parallel.sum_over_ranks(n_ev, dev)
    value = total_ev / ms / 1e3
    timed_region_s = float(np.sum(reps_ms)) / 1e3

    # ---- e2e: pinned host buffers -> H2D -> bin -> sample -> D2H, 3-stream pipeline ------------
    # The host side holds what the reference's loader reads from disk: raw 8-byte PSEE .dat Event2D records
    # (dat_events_tools.py:24) of the windows plus one record range per window; decode + binning + sampling
    # run on the GPU through the module API (AdaptiveRSNNEmbedding.forward_dat).
    from eas_snn_b200 import psee
    host_rec, host_rng = [], []
    for hb_np in host_np:
        x_, y_, t_, p_, off_ = hb_np
        host_rec.append(torch.from_numpy(psee.pack_records(x_, y_, t_, p_)).pin_memory())
        host_rng.append(torch.from_numpy(np.stack([off_[:-1], off_[1:]], axis=1).astype(np.int64)).pin_memory())
    nmax = max(r.shape[0] for r in host_rec)
    s_in, s_cmp, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    slots = []
    for _ in range(2):
        slots.append(dict(rec=torch.empty((nmax, 2), dtype=torch.int32, device=dev),
                          rng=torch.empty((BATCH, 2), dtype=torch.int64, device=dev),
                          host_out=torch.empty((TS, BATCH, 2, H, W), dtype=torch.float32).pin_memory(),
                          in_ready=torch.cuda.Event(), cmp_done=torch.cuda.Event(), out_done=torch.cuda.Event()))
    h2d_bytes = int(np.mean([r.numel() * 4 + g.numel() * 8 for r, g in zip(host_rec, host_rng)]))
    d2h_bytes = TS * BATCH * 2 * H * W * 4

    def e2e_loop(steps):
        n = 0
        for k in range(steps):
            sl, hr, hg = slots[k % 2], host_rec[k % NSETS], host_rng[k % NSETS]
            nrec = hr.shape[0]
            with torch.cuda.stream(s_in):
                s_in.wait_event(sl["cmp_done"])          # slot's previous compute finished reading inputs
                sl["rec"][:nrec].copy_(hr, non_blocking=True)
                sl["rng"].copy_(hg, non_blocking=True)
                sl["in_ready"].record(s_in)
            with torch.cuda.stream(s_cmp):
                s_cmp.wait_event(sl["in_ready"])
                # the frames this slot produced two steps ago have reached the host: their memory (freed just
                # below, when the slot's reference is replaced) may be reused by this step's allocations
                s_cmp.wait_event(sl["out_done"])
                with torch.no_grad():
                    frames = model.forward_dat(sl["rec"][:nrec], sl["rng"], H, W)   # public module API
                sl["frames"] = frames
                sl["cmp_done"].record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(sl["cmp_done"])
                sl["host_out"].copy_(frames, non_blocking=True)
                sl["out_done"].record(s_out)
            n += nrec
        return n

    e2e_loop(max(args.warmup, 10))       # long enough for the caching allocator to reach its steady state
    barrier()

    def timed_e2e(loop):
        t0 = time.perf_counter()
        n = loop(args.steps)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        barrier()
        return parallel.max_over_ranks(wall, dev), n

    w0, n_e2e = timed_e2e(e2e_loop)
    e2e_reps = int(min(200, max(1, np.ceil(MIN_TIMED_S * 1e3 / max(w0, 1e-3)))))
    if world > 1:
        e2e_reps = int(parallel.max_over_ranks(float(e2e_reps), dev))
    e2e_all = [w0] + [timed_e2e(e2e_loop)[0] for _ in range(e2e_reps - 1)]
    e2e_ms = float(np.median(e2e_all))
    e2e_val = parallel.sum_over_ranks(n_e2e, dev) / e2e_ms / 1e3
    checksum = float(slots[(args.steps - 1) % 2]["host_out"].abs().sum())

    # host <-> device ceiling: the SAME copies (sizes, pinned buffers, streams, all ranks at once) with no kernel in
    # between -- what the host memory / PCIe path of this box can feed; e2e is judged against it
    def copy_loop(steps):
        n = 0
        for k in range(steps):
            sl, hr, hg = slots[k % 2], host_rec[k % NSETS], host_rng[k % NSETS]
            with torch.cuda.stream(s_in):
                sl["rec"][:hr.shape[0]].copy_(hr, non_blocking=True)
                sl["rng"].copy_(hg, non_blocking=True)
            with torch.cuda.stream(s_out):
                sl["host_out"].copy_(sl["frames"], non_blocking=True)
            n += hr.shape[0]
        return n

    copy_loop(5)
    barrier()
    c_all = [timed_e2e(copy_loop) for _ in range(max(3, min(20, e2e_reps)))]
    ceil_ms = float(np.median([c[0] for c in c_all]))
    ceil_val = parallel.sum_over_ranks(c_all[0][1], dev) / ceil_ms / 1e3
    clk = clocks.stop()

    # ---- secondary metric: SYOLOX-M frames/s forward (T=3, 256x320): events -> detections ----
    frames = None
    if not args.no_backbone:
        from eas_snn_b200 import detector, fused

        def build_det(dw, mode, seed):
            torch.manual_seed(seed)
            d = detector.build_syolox(dw[0], dw[1], num_classes=2, T=3, embedding=model, use_spike=mode).to(dev).eval()
            for mod in d.modules():                # random init is dead: shift BN so layers fire
                if isinstance(mod, torch.nn.BatchNorm2d):
                    mod.bias.data.fill_(0.6)
            return d

        # e_yolox_m.py:13-14; use_spike full_spike = the configuration the README trains / evaluates SYOLOX-M with
        # (readme.md:136-160, event_yolox_base.py:207-211): spiking backbone AND spiking pyramid, ANN head on firing rates
        det = build_det((0.67, 0.75), "full_spike", 81)
        det_t = build_det((0.67, 0.75), True, 81)          # use_spike True (round-1 headline variant) for comparison
        bb = det.backbone.backbone

        def pad(fr):                               # multiples of 32 (event_yolox_base.py:556-559)
            return torch.nn.functional.pad(fr, (0, 320 - W, 0, 256 - H))

        def time_frames(fn):
            for w in range(3):
                out = fn(devb[w % NSETS])
            barrier()

            def once(fsteps):
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record()
                for k in range(fsteps):
                    o = fn(devb[k % NSETS])
                f1.record()
                barrier()
                return parallel.max_over_ranks(f0.elapsed_time(f1), dev) / fsteps, o

            t1, out = once(5)
            fsteps = int(min(400, max(5, np.ceil(MIN_TIMED_S * 1e3 / max(t1, 1e-3)))))
            if world > 1:
                fsteps = int(parallel.max_over_ranks(float(fsteps), dev))
            t, out = once(fsteps)
            return t, out, fsteps

        bb_ms, outs, _ = time_frames(lambda db: bb(pad(step(db))))          # spiking CSPDarknet only
        det_ms, pred, fsteps = time_frames(lambda db: det.detect_frames(pad(step(db))))   # + pyramid + head + decode
        dett_ms, pred_t, _ = time_frames(lambda db: det_t.detect_frames(pad(step(db))))
        det_t.set_ann_precision("fp16")            # the reduced-precision tier (reference: --fp16 evaluation)
        det16_ms, _, _ = time_frames(lambda db: det_t.detect_frames(pad(step(db))))
        det_t.set_ann_precision("fp32")
        n_spk = sum(1 for mod in det.modules() if isinstance(mod, fused.FusedConvBNPLIF)) + 1
        n_ann = sum(1 for mod in det.modules() if isinstance(mod, fused.AnnBaseConv)) - 1 + 6   # (- stem, + predictors)
        gflop_bb = 6.61 * 3 * BATCH                                          # M@256x320, per sample-step
        gflop_fs = 9.96 * 3 * BATCH                                          # whole PAFPN, spiking in full_spike
        frames = {"value": world * BATCH / det_ms * 1e3, "unit": "frames/s", "ms_per_batch": det_ms, "steps": fsteps,
                  "what": "events -> bin -> sampler -> SYOLOX-M forward, use_spike full_spike (T=3, 256x320): spiking "
                          "CSPDarknet + spiking pyramid (%d tcgen05 conv+BN+PLIF launches) -> time mean -> ANN YOLOX head "
                          "(%d tcgen05 conv launches, fp16 hi/lo split = fp32-equivalent) -> decoded predictions "
                          "[B, 1680, 7]; %d windows per GPU" % (n_spk, n_ann, BATCH),
                  "tensor": {"achieved": gflop_fs / det_ms, "unit": "TFLOP/s (1x conv FLOPs of backbone + pyramid; the "
                             "kernel runs 2 fp16 passes for fp32-equivalent weights)", "peak": 1394.4,
                             "frac": gflop_fs / det_ms / 1394.4},
                  "backbone_only": {"value": world * BATCH / bb_ms * 1e3, "ms_per_batch": bb_ms,
                                    "tensor": {"achieved": gflop_bb / bb_ms, "unit": "TFLOP/s (1x conv FLOPs; the "
                                               "kernel runs 2 fp16 passes for fp32-equivalent weights)",
                                               "peak": 1394.4, "frac": gflop_bb / bb_ms / 1394.4}},
                  "use_spike_true": {"value": world * BATCH / dett_ms * 1e3, "ms_per_batch": dett_ms,
                                     "note": "spiking backbone, ANN pyramid + head (three product terms per MMA slot): the "
                                             "variant the reference does not ship for SYOLOX-M",
                                     "pred_checksum": float(pred_t.float().abs().mean())},
                  "ann_fp16_activations": {"value": world * BATCH / det16_ms * 1e3, "ms_per_batch": det16_ms,
                                           "note": "use_spike True with fp16 activations in the ANN part (two product "
                                                   "terms): predictions within 1e-2 of the fp32 ones (tests); not the headline"},
                  "spike_rate": {k: round(float(v.float().mean()), 4) for k, v in outs.items()},
                  "pred_checksum": float(pred.float().abs().mean())}
        del det_t

        # the same path end to end: pinned host .dat records in (39 MB), decoded predictions out (3 MB) -- the frames never
        # leave the device, so this leg is bound by the detector, not by the host link like the sampler-only e2e
        pred_host = [torch.empty((BATCH, pred.shape[1], pred.shape[2]), dtype=torch.float32).pin_memory() for _ in range(2)]

        def frames_e2e_loop(steps):
            for k in range(steps):
                sl, hr, hg = slots[k % 2], host_rec[k % NSETS], host_rng[k % NSETS]
                nrec = hr.shape[0]
                with torch.cuda.stream(s_in):
                    s_in.wait_event(sl["cmp_done"])
                    sl["rec"][:nrec].copy_(hr, non_blocking=True)
                    sl["rng"].copy_(hg, non_blocking=True)
                    sl["in_ready"].record(s_in)
                with torch.cuda.stream(s_cmp):
                    s_cmp.wait_event(sl["in_ready"])
                    s_cmp.wait_event(sl["out_done"])
                    with torch.no_grad():
                        pr = det.detect_frames(pad(model.forward_dat(sl["rec"][:nrec], sl["rng"], H, W)))
                    sl["pred"] = pr
                    sl["cmp_done"].record(s_cmp)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(sl["cmp_done"])
                    pred_host[k % 2].copy_(pr, non_blocking=True)
                    sl["out_done"].record(s_out)
            return steps * BATCH

        frames_e2e_loop(4)
        barrier()
        fe_steps = int(min(400, max(5, np.ceil(MIN_TIMED_S * 1e3 / max(det_ms, 1e-3)))))
        if world > 1:
            fe_steps = int(parallel.max_over_ranks(float(fe_steps), dev))
        t0 = time.perf_counter()
        frames_e2e_loop(fe_steps)
        torch.cuda.synchronize()
        fe_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        fe_ms = parallel.max_over_ranks(fe_ms, dev) / fe_steps
        frames["e2e"] = {"value": world * BATCH / fe_ms * 1e3, "unit": "frames/s", "ms_per_batch": fe_ms, "steps": fe_steps,
                         "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(pred.numel() * 4),
                         "what": "pinned host .dat records + record ranges -> H2D -> forward_dat -> SYOLOX-M full_spike -> "
                                 "decoded predictions -> D2H into pinned host memory; 3 streams, 2 slots, wall clock",
                         "pred_checksum": float(pred_host[(fe_steps - 1) % 2].abs().mean())}

    # ---- secondary metric: 1Mpx inference (BASELINE config 3): RVT stacked histograms -> detections ------------
    mpx = None
    if not args.no_backbone:
        MB, MH, MW, NB10 = 16, 360, 640, 10
        gm = torch.Generator(device=dev).manual_seed(1234 + 3000 + rank)
        # uint8 [B*Tm, 2*10, 360, 640] (channel = polarity * 10 + bin), ~4 % occupied bins with counts 1..8
        occ = torch.rand((MB * TM, 2 * NB10, MH, MW), device=dev, generator=gm) < 0.04
        rep = (occ * torch.randint(1, 9, occ.shape, device=dev, generator=gm)).to(torch.uint8)
        del occ

        def mpx_step():
            ev = eas.rvt_event_sum(rep, NB10).view(MB, TM, 2, MH, MW)            # rvt_gen4.py:120-122 ('event_sum')
            with torch.no_grad():
                fr = model(ev)                                                  # Tm consecutive slices = sampler steps
            return det.detect_frames(torch.nn.functional.pad(fr, (0, 0, 0, 384 - MH)))   # 360x640 -> 384x640

        for _ in range(3):
            mp = mpx_step()
        barrier()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        msteps = 160                                   # ~6 ms each: a region of about a second
        m0.record()
        for _ in range(msteps):
            mp = mpx_step()
        m1.record()
        barrier()
        mms = parallel.max_over_ranks(m0.elapsed_time(m1), dev) / msteps
        mpx = {"value": world * MB / mms * 1e3, "unit": "frames/s", "ms_per_batch": mms, "steps": msteps,
               "what": "RVT-preprocessed uint8 [B*Tm, 20, 360, 640] stacked histograms -> event_sum -> adaptive "
                       "sampler (Tm=4 slices as steps, tensor-core kernel) -> zero pad to 384x640 -> whole SYOLOX-M "
                       "forward, use_spike full_spike (T=3) -> decoded predictions [B, 5040, 7]; %d windows per GPU, "
                       "input %d MB resident in HBM" % (MB, rep.numel() >> 20),
               "pred_checksum": float(mp.float().abs().mean())}
        del rep

    # ---- secondary metric: batch-1 latency (BASELINE config 1 on the GPU): SYOLOX-S, one window -> predictions ----
    latency = None
    if not args.no_backbone:
        from eas_snn_b200 import fused
        torch.manual_seed(83)
        det_s = detector.build_syolox(0.33, 0.50, num_classes=2, T=3, embedding=model).to(dev).eval()   # e_yolox_s.py:13-14
        for mod in det_s.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.bias.data.fill_(0.6)
        o1 = devb[0][4][:2]
        n1 = int(o1[-1])
        ev1 = tuple(a[:n1] for a in devb[0][:4])
        fr1 = pad(model(eas.bin_events(*ev1, o1, H, W, TM, dtype=torch.float32))).contiguous()
        graph = fused.GraphedForward(det_s.detect_frames, fr1)          # ~110 launches replayed by one host call
        hist1 = torch.empty((1, TM, 2, H, W), dtype=torch.float32, device=dev)

        def lat_step():
            with torch.no_grad():
                fr = model(eas.bin_events(*ev1, o1, H, W, TM, out=hist1))
            return graph(pad(fr))

        for _ in range(5):
            lat_step()
        barrier()
        ts = []
        for _ in range(30):
            l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0.record()
            lp = lat_step()
            l1.record()
            torch.cuda.synchronize()
            ts.append(l0.elapsed_time(l1))
        latency = {"value": float(np.median(ts)), "unit": "ms", "higher_is_better": False,
                   "what": "SYOLOX-S, batch 1, T=3: %d events of one 50 ms window -> bins -> sampler -> detector (CUDA "
                           "graph replay) -> decoded predictions [1, 1680, 7]; median of 30" % n1,
                   "pred_checksum": float(lp.float().abs().mean())}

    # ---- secondary metric: SYOLOX-S training step (BASELINE config 4), 8 windows per GPU --------------------
    train = None
    if not args.no_train:
        from eas_snn_b200 import fused
        TB = 8
        torch.manual_seed(82)
        t_emb = eas.AdaptiveRSNNEmbedding(**SAMPLER_KW).to(dev).train()
        t_bb = fused.SpikingCSPDarknet(0.33, 0.50, in_dim=2, T=3).to(dev).train()       # e_yolox_s.py:13-14
        # channels-last parameters: cuDNN then runs NHWC convolutions / BN end to end and the neurons work on those
        # buffers as they lie -- the NCHW<->NHWC conversion kernels were 10 % of the step
        t_bb = t_bb.to(memory_format=torch.channels_last)
        for mod in t_bb.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.bias.data.fill_(0.6)
        t_params = list(t_emb.parameters()) + list(t_bb.parameters())
        opt = torch.optim.Adam(t_params, lr=1e-4)
        t_off = devb[0][4][:TB + 1]
        n_t = int(t_off[-1])
        t_hist = eas.bin_events(devb[0][0][:n_t], devb[0][1][:n_t], devb[0][2][:n_t], devb[0][3][:n_t], t_off, H, W, TM,
                                dtype=torch.float32)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]

        def train_step(timed=False):
            opt.zero_grad(set_to_none=True)
            ev[0].record()
            fr = torch.nn.functional.pad(t_emb(t_hist), (0, 320 - W, 0, 256 - H))
            outs = t_bb(fr)
            loss = sum((v.mean() - 0.2) ** 2 for v in outs.values())     # proxy loss (the SimOTA head is out of scope)
            ev[1].record()
            loss.backward()
            ev[2].record()
            parallel.allreduce_gradients(t_params)                       # the reference's one collective (trainer.py:176)
            ev[3].record()
            opt.step()
            eas.reset_net(t_bb)                                          # trainer.py:115-117
            ev[4].record()
            return loss

        for _ in range(3):
            train_step()
        barrier()
        tsteps, parts = 5, np.zeros(4)
        t0 = time.perf_counter()
        for _ in range(tsteps):
            loss = train_step()
            torch.cuda.synchronize()
            parts += [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
        barrier()
        t_eager = parallel.max_over_ranks((time.perf_counter() - t0) * 1e3 / tsteps, dev)
        # the same step as two CUDA graphs around the all-reduce (fused.GraphedTrainStep): the host issues ~1500
        # launches per step in eager mode, about 1.5x the time the GPU needs for them
        opt_g = torch.optim.Adam(t_params, lr=1e-4, capturable=True, fused=True)
        loss_eager = float(loss.detach())
        del loss            # (the eager graph keeps the parameters' AccumulateGrad nodes bound to the default stream)
        import gc
        gc.collect()

        def t_loss(h):
            fr = torch.nn.functional.pad(t_emb(h), (0, 320 - W, 0, 256 - H))
            return sum((v.mean() - 0.2) ** 2 for v in t_bb(fr).values())

        gstep = fused.GraphedTrainStep(t_loss, [t_hist], t_params, opt_g, allreduce="flat",
                                       after=lambda: eas.reset_net(t_bb))
        for _ in range(3):
            gstep(t_hist)
        barrier()
        gsteps = 40
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(gsteps):
            loss = gstep(t_hist)
        g1.record()
        barrier()
        t_ms = parallel.max_over_ranks(g0.elapsed_time(g1) / gsteps, dev)
        train = {"value": world * TB / t_ms * 1e3, "unit": "samples/s", "ms_per_step": t_ms, "steps": gsteps,
                 "eager": {"ms_per_step": t_eager, "loss": loss_eager,
                           "parts_ms": dict(zip(("forward", "backward", "allreduce", "adam+reset"),
                                                (parts / tsteps).round(3).tolist()))},
                 "what": "SYOLOX-S training step, %d windows per GPU, T=3, 256x320, fp32, replayed as two CUDA graphs around "
                         "the gradient all-reduce (GraphedTrainStep; `eager` = the same step issued launch by launch): "
                         "sampler fwd + BPTT bwd (SAT surrogate + RPD) and 34 PLIF fwd/bwd on our kernels, conv / "
                         "batch-stat BN through cuDNN (channels-last), proxy loss on dark3-5 firing rates, NCCL gradient all-reduce "
                         "(%d params), Adam" % (TB, sum(p.numel() for p in t_params)),
                 "loss": float(loss.detach())}

    # ---- per-kernel durations (CUDA events on the launching stream) for the roofline -----------
    def time_call(fn, reps=10, warm=3, inner=16):
        """Median over `reps` of (CUDA-event time of `inner` back-to-back calls) / inner.  One call between two events
        measures the host's launch latency too (the GPU idles while Python issues a 60 us kernel); a queue of calls
        keeps the device busy, which is also how the step loop drives these kernels."""
        ts = []
        for r in range(warm + reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for i in range(inner):
                fn(r * inner + i)
            b.record()
            torch.cuda.synchronize()
            if r >= warm:
                ts.append(a.elapsed_time(b) / inner)
        return float(np.median(ts))

    # (four output buffers in turn, 150 MB: a single 37 MB buffer would simply stay in L2 between the calls)
    chist_rot = [chist_buf] + [eas.CompactHist.empty((BATCH, TM, 2, H, W), dev) for _ in range(NSETS - 1)]
    t_bin = time_call(lambda r: eas.bin_events(*devb[r % NSETS], H, W, TM, out=chist_rot[r % NSETS]))
    t_bin_f32 = time_call(lambda r: eas.bin_events(*devb[r % NSETS], H, W, TM, out=hist_buf))
    rec_dev = [(r.to(dev), g.to(dev)) for r, g in zip(host_rec, host_rng)]
    t_bin_dat = time_call(lambda r: psee.bin_dat(*rec_dev[r % NSETS], H, W, TM, out=chist_rot[r % NSETS]))
    del chist_rot
    hist_fixed = eas.bin_events(*devb[0], H, W, TM, dtype=torch.float32).clone()
    chist_fixed = eas.bin_events(*devb[0], H, W, TM, dtype=torch.uint8)
    with torch.no_grad():
        t_smp = time_call(lambda r: model(chist_fixed))
        t_smp_dense = time_call(lambda r: model(hist_fixed))
        model.algo = "fp32"
        t_smp_fp32 = time_call(lambda r: model(hist_fixed))
        model.algo = "auto"
    # stand-alone multi-step PLIF (a-4) through the C ABI: T = 3, the SYOLOX-M dark2 activation size
    import ctypes as C
    from eas_snn_b200 import _lib
    Lc = _lib.lib()
    pT, pN = 3, BATCH * 96 * 64 * 80
    px = torch.rand((pT, pN), device=dev) * 1.5
    ps, pg, pdx = torch.empty_like(px), torch.rand((pT, pN), device=dev), torch.empty_like(px)
    pw, pgw = torch.zeros((), device=dev), torch.zeros((), device=dev)
    pcfg = _lib.PlifCfg(T=pT, N=pN, v_threshold=1.0, hard_reset=0, v_reset=0.0, decay_input=0, detach_reset=0,
                        surrogate=0, alpha=2.0, dtype=_lib.EAS_F32)
    pwsb = Lc.eas_plif_bwd_ws_bytes(C.byref(pcfg))
    pws = torch.empty(pwsb, dtype=torch.uint8, device=dev)
    t_plif_f = time_call(lambda r: Lc.eas_plif_fwd(C.byref(pcfg), _lib.ptr(px), _lib.ptr(pw), None, _lib.ptr(ps), None,
                                                   _lib.stream_ptr()))
    t_plif_b = time_call(lambda r: Lc.eas_plif_bwd(C.byref(pcfg), _lib.ptr(px), _lib.ptr(pw), None, _lib.ptr(pg),
                                                   _lib.ptr(pdx), _lib.ptr(pgw), _lib.ptr(pws), pwsb, _lib.stream_ptr()))
    del px, ps, pg, pdx
    n_avg = float(np.mean([h.n for h in host]))
    bins = BATCH * TM * 2 * H * W
    bin_bytes = 13.0 * n_avg + 4.0 * bins                 # 13 B/event + 4 B/bin
    bin_bytes_touched = 5.0 * n_avg + 4.0 * bins          # t is only touched by the Tm+1 binary searches
    bin_bytes_compact = 5.0 * n_avg + 1.0 * bins          # what the compact call moves: 1 B per bin
    smp_launch_ms = t_smp / TM                                # one of the Tm step launches (+ 1/Tm of the weight pack)
    smp_bytes_launch = 8.0 * H * W * (TM + TS) * BATCH / TM   # SURVEY 8d, per launch (one of Tm steps)
    smp_flop_launch = 2400.0 * H * W * BATCH                  # 2400 FLOP per pixel-step
    fp32_peak = EAS_FP32_PEAK(sm_max_mhz)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    tensor_peak = float(peaks.get("bf16_tflops", 1590.0))     # burst figure: the kernel is timed alone
    # MMA model of the row-folded sampler kernel: 96 MMAs (M=128, K=16, N<=128) per 128-position tile
    # of 4 rows, 64 cycles each (scripts/umma_r4_probe.cu): the tensor pipe's own time per launch
    mma_floor_ms = tc2_mma_floor_cycles(BATCH, H, W) / (sm_max_mhz * 1e3)
    # DRAM traffic of the dominant kernel: read by key from the committed ncu summary of this kernel (never a literal)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r2_sampler_tc2_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), "profiles/r2_sampler_tc2_traffic.json: " + tj.get("what", "")
    with torch.no_grad():
        model.algo = "tensor_split"
        t_smp_v1 = time_call(lambda r: model(hist_fixed))
        model.algo = "auto"
    roofline = {
        "kernel": "sampler_tc2_step_kernel (dominant: %.0f%% of the step)" % (100.0 * t_smp / (t_smp + t_bin)),
        "bound": "tensor", "achieved": smp_flop_launch / smp_launch_ms / 1e9, "peak": tensor_peak, "unit": "TFLOP/s",
        "frac": smp_flop_launch / smp_launch_ms / 1e9 / tensor_peak,
        "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": "measured cuBLAS bf16 burst (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)",
        "launch_ms": smp_launch_ms, "launch_ms_dense_f32_input": t_smp_dense / TM,
        "note": "algorithmic FLOPs (2400 per pixel-step, SURVEY 8d) over the dense bf16 GEMM peak.  The convolutions have "
                "2/8 input and 4 output channels: on the tensor cores they run as block-Toeplitz GEMMs over the x axis "
                "(5 of 8 K positions used) x 4 output rows folded into N (5 of 8 input rows used per output row) x fp16 "
                "hi/lo planes of weights and hidden activations, i.e. ~20x the algorithmic FLOPs in MMA work; "
                "see mma_model for the tensor pipe's own time",
        "mma_model": {"floor_ms": mma_floor_ms, "frac": mma_floor_ms / smp_launch_ms,
                      "what": "96 tcgen05.mma (N = 32..128) per 2048-pixel tile x 64 cycles: at N = 128 an MMA runs at the "
                              "full math rate AND reads 8 KB of shared-memory operands per 64 cycles (128 B / clock)"},
        "hbm": {"achieved": smp_bytes_launch / smp_launch_ms / 1e6, "peak": peak_gbs, "unit": "GB/s",
                "frac": smp_bytes_launch / smp_launch_ms / 1e6 / peak_gbs,
                "note": "compulsory bytes only (240 FLOP/B: not the binding roofline)"},
        "first_tensor_kernel": {"launch_ms": t_smp_v1 / TM, "note": "sampler_tc_step_kernel (algo='tensor_split', round 1: 240 "
                                "MMAs of N = 32/48 per 2048 pixels, 18 B/element state round trip)"},
        "fp32_pipe_kernel": {"launch_ms": t_smp_fp32 / TM, "achieved": smp_flop_launch / (t_smp_fp32 / TM) / 1e9,
                             "peak": fp32_peak, "unit": "TFLOP/s", "frac": smp_flop_launch / (t_smp_fp32 / TM) / 1e9 / fp32_peak,
                             "note": "the FFMA2 kernel (algo='fp32', all other sampler configurations); peak = 148 SMs x 128 "
                                     "lanes x 2 x %.0f MHz" % sm_max_mhz},
        "others": {"bin_events (bounds + tiles)": {
            "bound": "hbm", "call_ms": t_bin, "achieved": bin_bytes / t_bin / 1e6, "peak": peak_gbs,
            "unit": "GB/s", "frac": bin_bytes / t_bin / 1e6 / peak_gbs,
            "achieved_bytes_moved": bin_bytes_compact / t_bin / 1e6,
            "frac_bytes_moved": bin_bytes_compact / t_bin / 1e6 / peak_gbs,
            "note": "the step's call: compact byte histogram.  achieved / frac = SURVEY 8d's algorithmic bytes (13 B per "
                    "event + 4 B per bin: the reference's dense histogram) over the call time; *_bytes_moved = what this "
                    "call really reads and writes (5 B per event + 1 B per bin)"},
            "bin_events, dense fp32 histogram": {
                "bound": "hbm", "call_ms": t_bin_f32, "achieved": bin_bytes / t_bin_f32 / 1e6, "peak": peak_gbs,
                "unit": "GB/s", "frac": bin_bytes / t_bin_f32 / 1e6 / peak_gbs,
                "achieved_bytes_touched": bin_bytes_touched / t_bin_f32 / 1e6,
                "note": "the public default (what the reference's loader produces); a plain memset of its 149 MB takes "
                        "0.046 ms on this GPU (scripts/write_bw_probe.py)"},
            "bin_dat (bounds + tiles on raw 8-byte records)": {
                "bound": "hbm", "call_ms": t_bin_dat, "achieved": (8.0 * n_avg + 4.0 * bins) / t_bin_dat / 1e6,
                "peak": peak_gbs, "unit": "GB/s", "frac": (8.0 * n_avg + 4.0 * bins) / t_bin_dat / 1e6 / peak_gbs,
                "note": "compact byte histogram; algorithmic bytes = 8 B/record + 4 B/bin"},
            "plif_fwd_kernel (f32, T=3)": {
                "bound": "hbm", "call_ms": t_plif_f, "achieved": 8.0 * pT * pN / t_plif_f / 1e6, "peak": peak_gbs,
                "unit": "GB/s", "frac": 8.0 * pT * pN / t_plif_f / 1e6 / peak_gbs, "note": "8 B per element-step"},
            "plif_bwd_kernel (f32, T=3, ATan)": {
                "bound": "hbm", "call_ms": t_plif_b, "achieved": 12.0 * pT * pN / t_plif_b / 1e6, "peak": peak_gbs,
                "unit": "GB/s", "frac": 12.0 * pT * pN / t_plif_b / 1e6 / peak_gbs,
                "note": "12 B per element-step (8 read + 4 written); the peak is the measured COPY bandwidth (half reads, "
                        "half writes), which a read-heavy kernel can slightly exceed"}},
    }

    # ---- BASELINE config 5: binning microbenchmark, one window of N events, Tm = 4 -----------------------------
    sweep = None
    if rank == 0 and not args.no_sweep:
        from eas_snn_b200 import synth
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > L2: every iteration starts cold
        sweep = {"what": "one 50 ms window of N events (uniform + 10 % hot cluster), Tm = 4, int32 histogram, strategy "
                         "auto; L2 flushed between iterations, CUDA events, median of 5; GB/s on SURVEY 8d's bytes "
                         "(13 B per event + 4 B per bin)", "rows": []}
        for (HH, WW) in ((240, 304), (720, 1280)):
            for N in (10 ** 5, 10 ** 6, 10 ** 7, 10 ** 8):
                rng = np.random.default_rng(1)
                x_, y_, t_, p_ = synth.make_window(rng, N, HH, WW)
                dd = [torch.from_numpy(a_).to(dev) for a_ in (x_, y_, t_, p_, np.array([0, N], np.int64))]
                del x_, y_, t_, p_
                outb = torch.empty((1, TM, 2, HH, WW), dtype=torch.int32, device=dev)
                ts_ = []
                for r in range(8):
                    flush.zero_()
                    a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a_.record()
                    eas.bin_events(*dd, HH, WW, TM, out=outb)
                    b_.record()
                    torch.cuda.synchronize()
                    if r >= 3:
                        ts_.append(a_.elapsed_time(b_))
                msb = float(np.median(ts_))
                byt = 13.0 * N + 4.0 * outb.numel()
                sweep["rows"].append({"frame": "%dx%d" % (HH, WW), "events": N, "ms": msb, "Mevents_per_s": N / msb / 1e3,
                                      "GBps": byt / msb / 1e6, "frac_hbm": byt / msb / 1e6 / peak_gbs,
                                      "checksum": int(outb.sum())})
                del dd, outb
        del flush

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(CONFIG),
            "config_notes": {"events_per_step_per_gpu": int(n_avg),
                             "parallelism": "sequence-sharded x%d, no collective" % world,
                             "l2": "inputs rotate over %d batches (%d MB events) + 150 MB histogram per step > 126 MB L2"
                                   % (NSETS, int(NSETS * n_avg * 13 / 1e6))},
            "timing": {"repeats": repeats, "timed_region_s": timed_region_s, "reported": "median repeat of exactly "
                       "%d steps" % args.steps, "min_ms": float(np.min(reps_ms)), "max_ms": float(np.max(reps_ms))},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": e2e_ms / args.steps, "repeats": e2e_reps,
                    "pipeline": "3 streams (H2D / compute / D2H), 2 slots",
                    "input": "raw 8-byte PSEE .dat records + one record range per window (what the reference's loader "
                             "reads from disk); decode, binning and sampling on the GPU via forward_dat",
                    "host_bw_ceiling": {"value": ceil_val, "unit": UNIT, "ms_per_step": ceil_ms / args.steps,
                                        "GBps_per_gpu": (h2d_bytes + d2h_bytes) / (ceil_ms / args.steps) / 1e6,
                                        "what": "the same pinned-host <-> device copies on the same streams, all ranks at "
                                                "once, no kernels: what this box's host memory / PCIe path can feed"},
                    "frac_of_host_ceiling": e2e_val / ceil_val,
                    "checksum": checksum, "numa": numa},
            # bin: bounds + histogram (2), sampler: weight pack (1) + Tm step launches + ONE cooperative fall-back
            # launch (idle unless the tensor-core kernel raised its flag)
            "gpu_launches": args.steps * repeats * (2 + 1 + TM + 1),
            "clocks": clk, "roofline": roofline}
    if sweep is not None:
        line["binning_sweep"] = sweep
    if frames is not None:
        line["frames"] = frames
    if mpx is not None:
        line["mpx"] = mpx
    if latency is not None:
        line["latency"] = latency
    if train is not None:
        line["train"] = train
    if rank == 0:
        os.sched_setaffinity(0, all_cpus)           # the CPU baseline gets every host core back
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
            if not args.no_backbone:
                line["cpu_baseline"]["frames"] = cpu_frames_baseline(budget_s=10.0, model="m")     # beside `frames`
                line["cpu_baseline"]["latency"] = cpu_frames_baseline(budget_s=8.0, model="s")     # beside `latency`
        _JSON_LINE.append(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def tc2_mma_floor_cycles(B: int, H: int, W: int, n_sm: int = 148) -> float:
    """Tensor-pipe time of one sampler step launch (the schedule of sampler_tc2.cu restated): every SM walks its share
    of (image, strip, row) space in segments; a segment of n rows has g = ceil(n/4) output row groups and costs
    ceil((g+1)*QPR/128) layer-1 tiles x 32 MMAs + ceil(g*QPR/128) layer-2 tiles x 64 MMAs, 64 cycles per MMA; the
    slowest SM counts."""
    max_tw = 4 * (31 - 2)
    ns = -(-W // max_tw)
    tw = (-(-W // ns) + 3) // 4 * 4
    qpr = tw // 4 + 2
    total = B * ns * H
    grid = min(n_sm, max(1, -(-total // 8)))
    worst = 0
    for c in range(grid):
        r, r_end, cyc = total * c // grid, total * (c + 1) // grid, 0
        while r < r_end:
            ya = r % H
            n = min(H - ya, r_end - r)
            g = -(-n // 4)
            cyc += (-(-((g + 1) * qpr) // 128) * 32 + -(-(g * qpr) // 128) * 64) * 64
            r += n
        worst = max(worst, cyc)
    return float(worst)


def EAS_FP32_PEAK(sm_mhz: float) -> float:
    return 148 * 128 * 2 * sm_mhz * 1e6 / 1e12


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-backbone", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly ONE line (the JSON): libraries that write to fd 1 (NCCL prints its version banner
    # there) are sent to stderr for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if _JSON_LINE:
        print(_JSON_LINE[0], flush=True)


if __name__ == "__main__":
    main()
