"""Generate the committed golden vectors by EXECUTING THE REFERENCE in the authoring container.

    python tests/golden/make_golden.py          # needs /root/reference (read-only)

Outputs (small, committed): tests/golden/binning.npz, sampler.npz, backbone.npz, psee.npz, detector.npz,
detector_full_spike.npz, detector_full_spike_v2.npz, count.npz, ablations.npz, voxel.npz, letterbox.npz.
The reference is imported in place through ``oracle/ref_loader.py``; nothing is copied from it.
``/root/reference`` does not exist on the GPU box, so tests only ever read the .npz files.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from eas_snn_b200 import synth  # noqa: E402


def struct_events(x, y, t, p):
    from yolox.utils.util import events_struct
    ev = np.zeros(len(x), dtype=events_struct)
    ev["x"], ev["y"], ev["t"], ev["p"] = x, y, t, p.astype(bool)
    return ev


def binning_cases():
    rng = np.random.default_rng(7)
    cases = []

    def add(name, x, y, t, p, H, W, Tm):
        cases.append(dict(name=name, x=np.asarray(x, np.int16), y=np.asarray(y, np.int16),
                          t=np.asarray(t, np.int64), p=np.asarray(p, np.uint8), H=H, W=W, Tm=Tm))

    x, y, t, p = synth.make_window(rng, 5000, 40, 48)
    add("uniform_40x48_tm4", x, y, t, p, 40, 48, 4)
    x, y, t, p = synth.make_window(rng, 3000, 40, 48, span_us=300)           # many duplicate timestamps
    add("dupes_40x48_tm5", x, y, t, p, 40, 48, 5)
    x, y, t, p = synth.make_window(rng, 4001, 33, 47, span_us=7919)          # odd H*W (scalar store path)
    add("odd_33x47_tm6", x, y, t, p, 33, 47, 6)
    add("single_event", [3], [2], [1000], [1], 8, 8, 4)                      # tw == 0 -> all empty
    add("three_events_tm4", [1, 2, 3], [1, 2, 3], [10, 11, 12], [0, 1, 0], 8, 8, 4)  # tw == 0
    add("same_timestamp", rng.integers(0, 8, 50), rng.integers(0, 8, 50), np.full(50, 77), rng.integers(0, 2, 50),
        8, 8, 4)
    add("exact_windows", np.arange(8) % 8, np.zeros(8), np.arange(8) * 10, np.arange(8) % 2, 8, 8, 4)
    add("offset_t0", rng.integers(0, 16, 400), rng.integers(0, 12, 400),
        np.sort(rng.integers(10**12, 10**12 + 50_000, 400)), rng.integers(0, 2, 400), 12, 16, 4)
    x, y, t, p = synth.make_window(rng, 20000, 240, 304)
    add("gen1_240x304_tm4", x, y, t, p, 240, 304, 4)
    # one pixel receiving > 65535 events in a micro-bin: exercises the 16-bit chunking of the tile kernel
    n = 150_000
    add("hot_pixel_overflow", np.full(n, 5), np.full(n, 6), np.sort(rng.integers(0, 1000, n)), np.ones(n), 16, 16, 2)
    return cases


def make_binning(gen1):
    out = {}
    cases = binning_cases()
    for c in cases:
        ds = ref_loader.make_ref_dataset(gen1, c["H"], c["W"], c["Tm"])
        ref = ds.agrregate(struct_events(c["x"], c["y"], c["t"], c["p"]), "micro_sum")
        assert ref.shape == (c["Tm"], 2, c["H"], c["W"]) and ref.dtype == np.float64
        assert np.all(ref == np.round(ref))
        n = c["name"]
        for k in ("x", "y", "t", "p"):
            out[f"{n}/{k}"] = c[k]
        out[f"{n}/dims"] = np.array([c["H"], c["W"], c["Tm"]], np.int64)
        out[f"{n}/hist"] = ref.astype(np.int32)
    # reference behaviour for "no events": zeros (gen1.py:356-358)
    ds = ref_loader.make_ref_dataset(gen1, 8, 8, 4)
    assert not ds.agrregate(None, "micro_sum").any()
    out["names"] = np.array([c["name"] for c in cases])
    np.savez_compressed(os.path.join(HERE, "binning.npz"), **out)
    print("binning.npz:", len(cases), "cases")


SAMPLER_CASES = [
    # name, B, H, W, Tm, Ts, ksize, depth, readout, vreset, spike_attach, write_zero, abs, rate
    ("published_sat_rpd", 2, 40, 48, 4, 1, 5, 2, "sum", 0, True, True, False, 1.5),
    ("ts2_avg_soft_abs", 2, 40, 48, 5, 2, 5, 2, "avg", None, False, False, True, 1.5),
    ("ts3_last_hard", 2, 40, 48, 6, 3, 5, 2, "last", 0, True, False, False, 1.5),
    ("depth1_k7_default", 2, 40, 48, 4, 1, 7, 1, "sum", 0, False, False, False, 1.0),
    ("depth1_k3_ts2", 1, 37, 70, 4, 2, 3, 1, "sum", None, True, True, False, 2.0),
    ("depth2_k3_ragged", 1, 19, 67, 3, 1, 3, 2, "sum", 0, True, True, False, 2.0),
    ("depth2_k7", 1, 24, 40, 4, 1, 7, 2, "sum", 0.25, False, False, False, 1.0),
]


def make_sampler(emb, act):
    from yolox.utils.util import warp_decay
    out = {}
    names = []
    for (name, B, H, W, Tm, Ts, k, depth, readout, vreset, sa, wz, ab, rate) in SAMPLER_CASES:
        torch.manual_seed(80)
        m = emb.AdaptiveRSNNEmbedding(kernel_size=k, in_channel=2, out_channel=2, readout=readout, split=False,
                                      write_zero=wz, abs=ab, depth=depth, nb_steps=Tm, vreset=vreset, thresh=1,
                                      spike_fn=act.Rectangle, decay=nn.Parameter(warp_decay(0.5)),
                                      embedding="arsnn", Ts=Ts, spike_attach=sa)
        g = torch.Generator().manual_seed(1234)
        x = torch.poisson(torch.full((B, Tm, 2, H, W), rate), generator=g)
        y = m(x)
        with torch.no_grad():  # fraction of pixels that fired at least once (t_last >= 0 at the end)
            _, rec = m(x, record=True)
        nz = float((rec[-1] >= 0).float().mean())
        assert 0.05 < nz < 0.95, (name, nz)
        wgt = torch.linspace(-1.0, 1.0, y.numel()).view_as(y)
        grads = torch.autograd.grad((y * wgt).sum(), list(m.parameters()))
        cfg = dict(B=B, H=H, W=W, Tm=Tm, Ts=Ts, ksize=k, depth=depth, readout=readout,
                   vreset=-1e30 if vreset is None else vreset, spike_attach=sa, write_zero=wz, abs=ab)
        out[f"{name}/cfg"] = np.array([repr(cfg)])
        out[f"{name}/x"] = x.numpy().astype(np.uint8)
        out[f"{name}/y"] = y.detach().numpy()
        for (pn, pv), gv in zip(m.named_parameters(), grads):
            out[f"{name}/param/{pn}"] = pv.detach().numpy()
            out[f"{name}/grad/{pn}"] = gv.numpy()
        names.append(name)
        print(" sampler", name, "fired %.3f" % nz, "out", tuple(y.shape))
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "sampler.npz"), **out)
    print("sampler.npz:", len(names), "cases")


def make_count(emb):
    """(f-4) The reference's SpikeCountEmbedding (embedding.py:9-24) on 5-D, 6-D and single-frame inputs."""
    m = emb.SpikeCountEmbedding(4)
    g = torch.Generator().manual_seed(21)
    x5 = torch.poisson(torch.full((3, 4, 2, 24, 32), 0.8), generator=g)
    x6 = torch.poisson(torch.full((2, 2, 4, 2, 24, 32), 0.8), generator=g)
    x4 = torch.poisson(torch.full((3, 2, 24, 32), 0.8), generator=g)
    out = {"x5": x5.numpy().astype(np.uint8), "y5": m(x5).numpy(), "x6": x6.numpy().astype(np.uint8),
           "y6": m(x6).numpy(), "x4": x4.numpy().astype(np.uint8), "y4": m(x4).numpy()}
    np.savez_compressed(os.path.join(HERE, "count.npz"), **out)
    print("count.npz:", {k: v.shape for k, v in out.items() if k.startswith("y")})


def make_ablations(emb, act):
    """(f-4) The reference's ablation embeddings -- ``LIFEmbedding`` ('snn', embedding.py:28-76) and
    ``SpikingEmbedding`` ('rsnn', :229-316) -- and the analysis outputs of the adaptive sampler (``record`` /
    ``v_record``, :198-199, 221-224) on Poisson micro-bin counts; parameters are stored with the outputs."""
    from yolox.utils.util import warp_decay
    out, names = {}, []
    cases = [("lif_d1_k5_sum", "lif", 5, 1, "sum", 0, 0.9), ("lif_d2_k3_last_soft", "lif", 3, 2, "last", None, 1.2),
             ("rsnn_d2_k5_sum", "rsnn", 5, 2, "sum", 0, 0.9), ("rsnn_d1_k7_last_relu", "rsnn", 7, 1, "last", 0.25, 1.1)]
    for (name, kind, k, depth, readout, vreset, rate) in cases:
        torch.manual_seed(80)
        kw = dict(nb_steps=4, vreset=vreset, thresh=1, spike_fn=act.Rectangle, decay=nn.Parameter(warp_decay(0.5)),
                  embedding=kind, Ts=1, spike_attach=True)
        if kind == "lif":
            m = emb.LIFEmbedding(kernel_size=k, in_channel=2, out_channel=2, readout=readout, depth=depth, **kw)
        else:
            m = emb.SpikingEmbedding(kernel_size=k, in_channel=2, out_channel=2, readout=readout,
                                     relu=name.endswith("relu"), depth=depth, **kw)
        for mm in m.modules():                      # non-zero biases (the default init leaves them tiny)
            if isinstance(mm, nn.Conv2d):
                nn.init.uniform_(mm.bias, -0.3, 0.3)
        g = torch.Generator().manual_seed(77)
        x = torch.poisson(torch.full((2, 4, 2, 40, 48), rate), generator=g)
        with torch.no_grad():
            y = m(x)
        assert y.shape == (2, 2, 40, 48) and float((y != 0).float().mean()) > 0.1
        out[f"{name}/x"] = x.numpy().astype(np.uint8)
        out[f"{name}/y"] = y.numpy()
        out[f"{name}/cfg"] = np.array([repr(dict(kind=kind, ksize=k, depth=depth, readout=readout,
                                                 vreset=-1e30 if vreset is None else vreset, relu=name.endswith("relu")))])
        for pn, pv in m.state_dict().items():
            out[f"{name}/sd/{pn}"] = pv.detach().numpy()
        names.append(name)
        print(" ablation", name, "out", tuple(y.shape), "mean |y| %.3f" % float(y.abs().mean()))
    # record / v_record of the adaptive sampler (two configurations: Ts = 1 with the early break never taken, Ts = 2)
    for name, Ts, rate in (("record_ts1", 1, 1.0), ("record_ts2", 2, 1.4)):
        torch.manual_seed(80)
        m = emb.AdaptiveRSNNEmbedding(kernel_size=5, in_channel=2, out_channel=2, readout="sum", split=False,
                                      write_zero=True, abs=False, depth=2, nb_steps=5, vreset=0, thresh=1,
                                      spike_fn=act.Rectangle, decay=nn.Parameter(warp_decay(0.5)), embedding="arsnn",
                                      Ts=Ts, spike_attach=True)
        g = torch.Generator().manual_seed(78)
        x = torch.poisson(torch.full((2, 5, 2, 40, 48), rate), generator=g)
        with torch.no_grad():
            y, rec = m(x, record=True)
            y2, vrec = m(x, v_record=True)
        assert torch.equal(y, y2)
        out[f"{name}/x"] = x.numpy().astype(np.uint8)
        out[f"{name}/y"] = y.numpy()
        out[f"{name}/record"] = rec.numpy().astype(np.int8)
        out[f"{name}/v_record"] = vrec.numpy()
        for pn, pv in m.state_dict().items():
            out[f"{name}/sd/{pn}"] = pv.detach().numpy()
        print(" record", name, "steps", rec.shape[0], "fired %.3f" % float((rec[-1] >= 0).float().mean()),
              "v_record", tuple(vrec.shape))
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "ablations.npz"), **out)


def make_voxel():
    """(f-4) ``to_voxel_grid_numpy`` (yolox/utils/event_reps.py:30-89) on synthetic windows: with the reference's own
    ``events_struct`` dtype (bool polarity: its in-place ``pols[pols == 0] = -1`` then stores True and every event weighs
    +1) and with a signed int8 polarity field (+1 / -1 as its comment intends).  Inputs come from seeds."""
    import importlib
    reps = importlib.import_module("yolox.utils.event_reps")
    out = {}
    for i, (n, H, W, nb) in enumerate(((3000, 40, 48, 10), (20000, 64, 80, 5), (1, 24, 32, 4), (2, 24, 32, 3))):
        rng = np.random.default_rng(500 + i)
        x, y, t, p = synth.make_window(rng, n, H, W)
        for tag, pdt in (("bool", bool), ("int8", np.int8)):
            ev = np.zeros(n, dtype=[("x", np.int16), ("y", np.int16), ("t", np.int64), ("p", pdt)])
            ev["x"], ev["y"], ev["t"], ev["p"] = x, y, t, p
            if n >= 2 and t[-1] > t[0]:
                vg = reps.to_voxel_grid_numpy(ev, (W, H, 2), n_time_bins=nb)
                out["%d/%s" % (i, tag)] = vg.astype(np.float64)
        out["%d/cfg" % i] = np.array([n, H, W, nb])
        print(" voxel", i, (n, H, W, nb), [k for k in out if k.startswith("%d/" % i)])
    np.savez_compressed(os.path.join(HERE, "voxel.npz"), **out)


LETTERBOX_CASES = [  # ih, iw, h, w, center, letterbox
    (240, 304, 640, 640, False, True),     # the reference's default input_size
    (240, 304, 256, 320, False, True),
    (240, 304, 512, 640, True, True),
    (64, 96, 48, 72, True, True),          # down-scaling
    (37, 53, 128, 96, False, True),        # odd source size
    (240, 304, 320, 320, False, False),    # no letterbox: stretched
    (40, 48, 40, 48, False, True),         # identity
]


def make_letterbox(gen1):
    """(f-3) The reference's own GEN1Dataset.get_random_data(random=False) (gen1.py:433-483; cv2.resize INTER_LINEAR
    per micro-frame + paste into a zero canvas) on synthetic count frames.  Inputs are regenerated from the seed by the
    tests; the fixture keeps a strided sample of every output plus its checksum."""
    out = {}
    for i, (ih, iw, h, w, center, lb) in enumerate(LETTERBOX_CASES):
        rng = np.random.default_rng(100 + i)
        fr = rng.poisson(0.7, (3, 2, ih, iw)).astype(np.float64)
        ds = object.__new__(gen1.GEN1Dataset)
        ds.letterbox_image = lb
        ref, _ = ds.get_random_data(fr, np.zeros((0, 5)), (h, w), random=False, center=center)
        assert ref.shape == (3, 2, h, w)
        out["%d/cfg" % i] = np.array([ih, iw, h, w, int(center), int(lb)], np.int64)
        out["%d/sample" % i] = ref[:, :, ::3, ::5].astype(np.float64)
        out["%d/sum" % i] = np.array([ref.sum(), np.abs(ref).max(), float((ref != 0).mean())])
    np.savez_compressed(os.path.join(HERE, "letterbox.npz"), **out)
    print("letterbox.npz:", len(LETTERBOX_CASES), "cases")


def make_backbone():
    """Tiny-width spiking CSPDarknet built by the REFERENCE code (darknet.py + utils_snn.py) through
    the spikingjelly shim; BN calibrated so that every stage fires."""
    exp, model = ref_loader.load_full_model("e-yolox-s", ["T", 3, "embedding", "arsnn", "embedding_depth", 2,
                                                          "embedding_ksize", 5, "spike_attach", True,
                                                          "write_zero", True, "spike_fn", "atan",
                                                          "use_spike", True, "num_classes", 2])
    from yolox.models.darknet import CSPDarknet
    from yolox.utils.utils_snn import convert_to_spiking
    from spikingjelly.activation_based import surrogate, functional
    from oracle.backbone import calibrate_bn
    torch.manual_seed(80)
    bb = CSPDarknet(0.33, 0.125, in_dim=2, act="silu")
    bb = convert_to_spiking(bb, surrogate.ATan(2.0))
    for m in bb.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.eps, m.momentum = 1e-3, 0.03
    g = torch.Generator().manual_seed(5)
    frame = torch.rand((1, 2, 2, 64, 96), generator=g) * 3.0
    x = frame.expand(3, -1, -1, -1, -1).contiguous()     # Ts == 1 broadcast to T == 3 (spiking_yolox.py:54-55)
    calibrate_bn(bb, x, seed=3)
    # perturb the PLIF decay parameter so sigmoid(w) != 0.5 is exercised
    for i, m in enumerate(mm for mm in bb.modules() if hasattr(mm, "w") and isinstance(mm.w, nn.Parameter)):
        m.w.data.fill_(0.3 * ((i % 5) - 2))
    with torch.no_grad():
        outs = bb(x)
        functional.reset_net(bb)
    out = {"x": x.numpy()}
    for k, v in bb.state_dict().items():
        out["sd/" + k] = v.numpy()
    for k, v in outs.items():
        rate = float(v.mean())
        assert 0.01 < rate < 0.9, (k, rate)
        out["out/" + k] = v.numpy().astype(np.uint8)
        print(" backbone", k, tuple(v.shape), "rate %.3f" % rate)
    np.savez_compressed(os.path.join(HERE, "backbone.npz"), **out)
    print("backbone.npz written, params:", sum(p.numel() for p in bb.parameters()))


def make_detector(use_spike=True):
    """(f-2) The reference's whole model -- ``EventExp.get_model()`` -- at a tiny width, run unmodified through the
    spikingjelly shim: micro-bin histograms in, decoded predictions out, plus the pyramid features, the sampler
    frames and the reference's own ``postprocess`` (boxes.py:33-77) on those predictions.
      use_spike True            : SpikingYOLOX(SpikingYOLOPAFPN, YOLOXHead, ...)               -> detector.npz
      use_spike 'full_spike'    : SpikingYOLOX(convert_to_spiking(YOLOPAFPN), SpikingYOLOXHead(full_spike=False))
                                  (event_yolox_base.py:207-211; what readme.md:136-160 trains / evaluates) -> detector_full_spike.npz
      use_spike 'full_spike_v2' : ... SpikingYOLOXHead(full_spike=True): spiking towers, time mean of the predictor
                                  outputs (spiking_yolo_head.py:125-127, 159-178)                -> detector_full_spike_v2.npz"""
    tag = "" if use_spike is True else "_" + use_spike
    exp, model = ref_loader.load_full_model("e-yolox-s", ["T", 3, "embedding", "arsnn", "embedding_depth", 2,
                                                          "embedding_ksize", 5, "spike_attach", True,
                                                          "write_zero", True, "spike_fn", "atan",
                                                          "use_spike", use_spike, "num_classes", 2,
                                                          "width", 0.125, "depth", 0.33])
    from spikingjelly.activation_based import functional
    from oracle.backbone import calibrate_bn
    from yolox.utils.boxes import postprocess
    torch.manual_seed(80)
    for part in (model.backbone, model.head):   # deterministic weights (get_model consumed an unknown amount of RNG)
        for m in part.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, a=5 ** 0.5)
    model.embedding._init_weight()
    for m in model.embedding.modules():          # _init_weight leaves the biases alone
        if isinstance(m, nn.Conv2d):
            bound = 1.0 / (m.weight[0].numel() ** 0.5)
            nn.init.uniform_(m.bias, -bound, bound)
    B, Tm, H, W = 2, exp.Tm, 64, 96
    g = torch.Generator().manual_seed(11)
    hist = torch.poisson(torch.full((B, Tm, 2, H, W), 0.6), generator=g)
    with torch.no_grad():
        frames = model.embedding(hist)                                   # [Ts=1, B, 2, H, W]
    nz = float((frames != 0).float().mean())
    print(" detector%s sampler frames" % tag, tuple(frames.shape), "non-zero %.3f" % nz)
    assert frames.shape == (1, B, 2, H, W) and 0.01 < nz < 0.9
    x = frames.expand(3, -1, -1, -1, -1).contiguous()
    calibrate_bn(model.backbone, x, seed=3)                              # every BN of the backbone + pyramid
    gg = torch.Generator().manual_seed(4)
    head_bns = [m for m in model.head.modules() if isinstance(m, nn.BatchNorm2d)]
    for m in head_bns:                                                   # head BNs: non-trivial affine parameters
        m.weight.data = torch.empty_like(m.weight).uniform_(0.8, 1.2, generator=gg)
        m.bias.data = torch.empty_like(m.bias).normal_(0.0, 0.1, generator=gg)
        if use_spike != "full_spike_v2":                                 # ANN head: any running statistics will do
            m.running_mean.data = torch.empty_like(m.running_mean).normal_(0.0, 0.2, generator=gg)
            m.running_var.data = torch.empty_like(m.running_var).uniform_(0.5, 1.5, generator=gg)
    if use_spike == "full_spike_v2":
        # spiking towers: running statistics calibrated on their own inputs (random ones leave the towers dead or
        # saturated -- the vacuous-parity hazard of SURVEY 7.8); two train-mode passes with momentum 1
        old = [m.momentum for m in head_bns]
        for m in head_bns:
            m.momentum = 1.0
        model.eval()
        with torch.no_grad():
            feats = model.backbone(x)
            functional.reset_net(model)
            model.head.train()
            for _ in range(2):
                for k, f in enumerate(feats):
                    y = model.head.stems[k](f)
                    model.head.cls_convs[k](y)
                    model.head.reg_convs[k](y)
                functional.reset_net(model)
        for m, o in zip(head_bns, old):
            m.momentum = o
    preds_convs = [m for ml in (model.head.cls_preds, model.head.reg_preds, model.head.obj_preds) for mm in ml
                   for m in mm.modules() if isinstance(m, nn.Conv2d)]
    for conv in preds_convs:
        # (the 1e-2 prior of initialize_biases would put every score below any useful threshold)
        conv.bias.data = torch.empty_like(conv.bias).normal_(0.0, 0.7, generator=gg)
        conv.weight.data *= 0.3
    for i, m in enumerate(mm for mm in model.modules() if hasattr(mm, "w") and isinstance(mm.w, nn.Parameter)):
        m.w.data.fill_(0.3 * ((i % 5) - 2))
    model.eval()
    with torch.no_grad():
        pyramid = model.backbone(x)
        functional.reset_net(model)
        pred = model(hist)                                               # [B, A, 7] decoded
        functional.reset_net(model)
        model.head.decode_in_inference = False
        raw = model(hist)
        functional.reset_net(model)
        dets = postprocess(pred.clone(), 2, conf_thre=0.3, nms_thre=0.45)
    A = sum((H // s) * (W // s) for s in (8, 16, 32))
    assert pred.shape == (B, A, 7), pred.shape
    out = {"hist": hist.numpy().astype(np.uint8), "frames": frames.numpy(), "pred": pred.numpy(), "raw": raw.numpy()}
    for k, v in model.state_dict().items():
        out["sd/" + k] = v.numpy()
    for i, f in enumerate(pyramid):
        if f.dim() == 5:                                                 # spiking pyramid: [T, B, C, H, W] spikes
            rate = float(f.mean())
            assert 0.01 < rate < 0.9, ("pyramid level %d is dead or saturated" % i, rate)
            assert bool(((f == 0) | (f == 1)).all())
            out["pyramid/%d" % i] = f.numpy().astype(np.uint8)
            print(" detector%s pyramid" % tag, i, tuple(f.shape), "rate %.3f" % rate)
        else:
            out["pyramid/%d" % i] = f.numpy()
            print(" detector%s pyramid" % tag, i, tuple(f.shape), "mean |x| %.3f" % float(f.abs().mean()))
    if use_spike == "full_spike_v2":                                     # firing rates inside the spiking towers
        with torch.no_grad():
            for k, f in enumerate(pyramid):
                y = model.head.stems[k](f)
                c, r = model.head.cls_convs[k](y), model.head.reg_convs[k](y)
                for nm, v in (("stem", y), ("cls", c), ("reg", r)):
                    rate = float(v.mean())
                    assert 0.005 < rate < 0.95, (nm, k, rate)
                out["tower/%d/cls" % k] = c.numpy().astype(np.uint8)
                out["tower/%d/reg" % k] = r.numpy().astype(np.uint8)
            functional.reset_net(model)
    n_det = 0
    for i, d in enumerate(dets):
        out["dets/%d" % i] = np.zeros((0, 7), np.float32) if d is None else d.numpy()
        n_det += 0 if d is None else len(d)
    assert n_det >= 4, "the fixture should produce a few detections (got %d)" % n_det
    out["meta"] = np.array([repr(dict(depth=0.33, width=0.125, num_classes=2, T=3, Tm=Tm, Ts=1, ksize=5,
                                      emb_depth=2, readout=exp.readout, vreset=exp.reset, thresh=exp.thresh,
                                      spike_attach=True, write_zero=True, abs=bool(exp.abs), alpha=float(exp.alpha),
                                      conf_thre=0.3, nms_thre=0.45, use_spike=use_spike,
                                      spikingjelly=ref_loader.spikingjelly_origin()))])
    np.savez_compressed(os.path.join(HERE, "detector%s.npz" % tag), **out)
    print("detector%s.npz written: pred" % tag, tuple(pred.shape), "detections", n_det,
          "obj range %.3f..%.3f" % (float(pred[..., 4].min()), float(pred[..., 4].max())))


def make_psee(gen1):
    """PSEE .dat path: synthetic recordings written in the reference's format (records by its own
    ``write_event_buffer``; the header by hand because the reference's ``write_header`` references an
    undefined name), searched by the reference's ``GEN1Dataset.search_events`` -> ``PSEELoader`` and
    aggregated by ``agrregate('micro_sum')``.  The fixture keeps outputs only; inputs come from seeds."""
    import importlib
    import tempfile
    dat_tools = importlib.import_module("yolox.utils.psee_loader.io.dat_events_tools")
    # numpy >= 2 refuses `python_int // np.uint8(8)` in PSEELoader.__init__ (the reference targets numpy 1.x):
    # hand it the same header values as Python ints; nothing else of the reference is touched
    _parse = dat_tools.parse_header
    dat_tools.parse_header = lambda f: tuple(int(v) if isinstance(v, np.integer) else v for v in _parse(f))
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, (kw, window, num_slice, Tm) in synth.DAT_CASES.items():
            x, y, t, p = synth.dat_stream(**kw)
            H, W = kw["H"], kw["W"]
            path = os.path.join(tmp, name + "_td.dat")
            with open(path, "w") as f:
                f.write("%% Data file containing Event2D events.\n%% Version 2\n%% Height %d\n%% Width %d\n" % (H, W))
                np.array([0, 8], dtype=np.uint8).tofile(f)
                f.flush()
                buf = np.zeros(len(t), dtype=[("x", "u2"), ("y", "u2"), ("p", "u1"), ("t", "u4")])
                buf["x"], buf["y"], buf["p"], buf["t"] = x, y, p, t
                dat_tools.write_event_buffer(f, buf)
            ds = object.__new__(gen1.GEN1Dataset)
            ds.img_size = (H, W)
            ds.files = [os.path.join(tmp, name + "_bbox.npy")]
            ds.slice_policy = "fix_t"
            ds.slice_args = {"window": window, "num_slice": num_slice, "micro_slice": Tm, "aggregation": "micro_sum"}
            ts = synth.dat_label_times(name, t, window)
            first = np.zeros(len(ts), np.int64)      # index of the first returned event in the recording
            count = np.zeros(len(ts), np.int64)
            hists = np.zeros((len(ts), Tm, 2, H, W), np.int32)
            for k, stamp in enumerate(ts):
                ev = ds.search_events(0, int(stamp))
                count[k] = len(ev)
                if len(ev):
                    # locate the slice in the recording: (t, x, y, p) of its first event + its length
                    cand = np.flatnonzero(t == ev["t"][0])
                    hit = [c for c in cand if c + len(ev) <= len(t) and np.array_equal(t[c:c + len(ev)], ev["t"])
                           and np.array_equal(x[c:c + len(ev)], ev["x"]) and np.array_equal(y[c:c + len(ev)], ev["y"])]
                    assert len(hit) >= 1
                    first[k] = hit[0]
                fr = ds.agrregate(ev, "micro_sum")
                assert fr.shape == (Tm, 2, H, W) and np.all(fr == np.round(fr))
                hists[k] = fr.astype(np.int32)
            out[f"{name}/ts"] = ts
            out[f"{name}/first"] = first
            out[f"{name}/count"] = count
            out[f"{name}/hist"] = hists
            print("psee", name, "labels", len(ts), "empty windows", int((count == 0).sum()),
                  "events/window max", int(count.max()))
    out["names"] = np.array(list(synth.DAT_CASES))
    np.savez_compressed(os.path.join(HERE, "psee.npz"), **out)


if __name__ == "__main__":
    assert ref_loader.available(), "reference tree not found"
    torch.set_num_threads(8)
    emb, act, gen1 = ref_loader.load_hot_modules()
    if len(sys.argv) > 1 and sys.argv[1] == "psee":
        make_psee(gen1)
        print("psee.npz", os.path.getsize(os.path.join(HERE, "psee.npz")) // 1024, "KiB")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "f4":
        make_ablations(emb, act)
        make_voxel()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "letterbox":
        make_letterbox(gen1)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "detector":
        make_detector()
        print("detector.npz", os.path.getsize(os.path.join(HERE, "detector.npz")) // 1024, "KiB")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "full_spike":
        make_detector("full_spike")
        make_detector("full_spike_v2")
        sys.exit(0)
    make_binning(gen1)
    make_sampler(emb, act)
    make_count(emb)
    make_ablations(emb, act)
    make_voxel()
    make_letterbox(gen1)
    make_psee(gen1)          # (before the two below: load_full_model re-imports yolox without the package stubs)
    make_backbone()
    make_detector()
    make_detector("full_spike")
    make_detector("full_spike_v2")
    for f in ("binning.npz", "sampler.npz", "count.npz", "letterbox.npz", "backbone.npz", "psee.npz", "detector.npz",
              "detector_full_spike.npz", "detector_full_spike_v2.npz", "ablations.npz", "voxel.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
