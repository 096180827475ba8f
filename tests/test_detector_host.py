"""CPU: host-side logic of the whole-detector modules (no kernels run): the reference checkpoint layout loads key for
key, the product refuses to run without CUDA, postprocess reproduces the reference's detections on its predictions."""
import numpy as np
import pytest
import torch

import eas_snn_b200 as eas
from eas_snn_b200 import _lib, detector
from helpers import detector_case, detector_sampler_kwargs, load_golden


def _net():
    z = load_golden("detector")
    meta, sd, hist = detector_case(z)
    emb = eas.AdaptiveRSNNEmbedding(**detector_sampler_kwargs(meta))
    net = detector.build_syolox(meta["depth"], meta["width"], meta["num_classes"], meta["T"], embedding=emb,
                                spike_fn=eas.ATan(meta["alpha"]))
    return z, meta, sd, hist, net


def test_reference_state_dict_loads_key_for_key():
    z, meta, sd, hist, net = _net()
    assert set(net.state_dict()) == set(sd)                       # same names, nothing extra, nothing missing
    net.load_state_dict(sd, strict=True)
    for k, v in net.state_dict().items():
        assert v.shape == sd[k].shape, k
    # the names the reference's optimizer grouping and checkpoints rely on (event_yolox_base.py:389-403)
    for k in ("embedding.gate_conv.0.weight", "backbone.backbone.stem.0.conv.conv.weight",
              "backbone.backbone.dark2.0.act.w", "backbone.lateral_conv0.bn.running_mean",
              "backbone.C3_n4.m.0.conv2.conv.weight", "head.stems.2.conv.weight", "head.cls_preds.0.bias",
              "head.obj_preds.1.weight"):
        assert k in sd, k
    assert sum(eas.is_spiking_neuron(m) for m in net.modules()) == sum(k.endswith("act.w") for k in sd)


@pytest.mark.parametrize("mode", ["full_spike", "full_spike_v2"])
def test_full_spike_state_dicts_load_key_for_key(mode):
    """The reference's ``use_spike full_spike / full_spike_v2`` checkpoints (event_yolox_base.py:207-211): converted
    pyramid (``lateral_conv0.conv.0.weight``, ``...act.w``) and, for v2, converted head (``stems.0.conv.0.weight``,
    ``cls_preds.0.0.bias``) -- same names, nothing extra, nothing missing."""
    z = load_golden("detector_" + mode)
    meta, sd, hist = detector_case(z)
    emb = eas.AdaptiveRSNNEmbedding(**detector_sampler_kwargs(meta))
    net = detector.build_syolox(meta["depth"], meta["width"], meta["num_classes"], meta["T"], embedding=emb,
                                spike_fn=eas.ATan(meta["alpha"]), use_spike=mode)
    assert set(net.state_dict()) == set(sd)
    net.load_state_dict(sd, strict=True)
    want = ["backbone.lateral_conv0.conv.0.weight", "backbone.C3_p4.m.0.conv2.act.w", "backbone.bu_conv1.bn.running_var"]
    want += ["head.stems.0.conv.0.weight", "head.cls_convs.1.0.act.w", "head.cls_preds.0.0.bias"] if mode.endswith("v2") \
        else ["head.stems.0.conv.weight", "head.cls_preds.0.bias"]
    for k in want:
        assert k in sd, k
    n_plif = sum(eas.is_spiking_neuron(m) for m in net.modules())
    assert n_plif == sum(k.endswith("act.w") for k in sd)
    # the M-sized models have 82 / 97 neurons; the S-sized fixture 58 / 73 (34 backbone + 24 pyramid [+ 15 head])
    assert n_plif == (73 if mode.endswith("v2") else 58), n_plif


def test_syolox_m_full_spike_neuron_counts():
    for mode, want in (("full_spike", 82), ("full_spike_v2", 97)):       # SURVEY 8a-4 [probe]
        net = detector.build_syolox(0.67, 0.75, 2, 3, use_spike=mode)
        assert sum(eas.is_spiking_neuron(m) for m in net.modules()) == want, mode


def test_syolox_s_and_m_parameter_counts_match_the_reference_readme():
    # readme.md model table: e-yolox-s 8.94 M, e-yolox-m 25.28 M parameters (SURVEY 8c probe)
    for (d, w, want) in ((0.33, 0.50, 8_938_167), (0.67, 0.75, 25_281_000)):
        emb = eas.AdaptiveRSNNEmbedding(kernel_size=5, depth=2, nb_steps=4, thresh=1, vreset=0)
        n = sum(p.numel() for p in detector.build_syolox(d, w, 2, 3, embedding=emb).parameters())
        assert abs(n - want) <= 1000, (d, w, n)


def test_detector_refuses_cpu_tensors_and_training_mode():
    z, meta, sd, hist, net = _net()
    net.eval()
    with pytest.raises(_lib.EasError):
        net(hist)                                  # no CPU fallback anywhere on the path
    net.train()
    with pytest.raises(RuntimeError):
        net(hist)


def test_pad_frames_and_postprocess_match_the_reference_detections():
    z, meta, *_ = _net()
    f = torch.arange(2 * 3 * 5 * 7, dtype=torch.float32).view(1, 2, 3, 5, 7)
    p = detector.pad_frames(f)
    assert p.shape == (1, 2, 3, 32, 32) and torch.equal(p[..., :5, :7], f) and float(p.sum()) == float(f.sum())
    assert detector.pad_frames(p) is p
    pred = torch.from_numpy(z["pred"])
    keep = pred.clone()
    dets = detector.postprocess(pred, meta["num_classes"], meta["conf_thre"], meta["nms_thre"])
    assert torch.equal(pred, keep)                 # unlike boxes.py:33-41 the input is left alone
    n = 0
    for i, d in enumerate(dets):
        want = torch.from_numpy(z["dets/%d" % i])
        got = torch.zeros((0, 7)) if d is None else d
        assert got.shape == want.shape and torch.allclose(got, want, rtol=0, atol=1e-5), i
        n += len(want)
    assert n >= 4
    # class-agnostic branch and the empty case
    ag = detector.postprocess(pred, meta["num_classes"], meta["conf_thre"], meta["nms_thre"], class_agnostic=True)
    assert all(a is not None and len(a) <= len(d) for a, d in zip(ag, dets))
    assert detector.postprocess(pred, meta["num_classes"], conf_thre=2.0) == [None, None]


def test_convert_to_spiking_gives_the_reference_checkpoint_layout():
    """``fused.convert_to_spiking`` (the seam INTEGRATION.md Level 1 names; utils_snn.py:16-58) on a plain ANN
    CSPDarknet: every BaseConv becomes the fused conv+BN+PLIF layer, Focus is wrapped whole and stays ANN, lone
    MaxPool2d are wrapped -- and the result carries exactly the keys of the REFERENCE's converted backbone
    (tests/golden/backbone.npz, produced by the reference's own convert_to_spiking)."""
    from eas_snn_b200 import fused
    from helpers import AnnCSPDarknet
    z = load_golden("backbone")
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    net = fused.convert_to_spiking(AnnCSPDarknet(0.33, 0.125, in_dim=2), eas.ATan(2.0))
    assert set(net.state_dict()) == set(sd)
    net.load_state_dict(sd, strict=True)
    assert isinstance(net.stem, fused.SeqToANNContainer) and type(net.stem[0]).__name__ == "Focus"
    assert isinstance(net.stem[0].conv.act, torch.nn.SiLU)                  # the stem stays ANN (utils_snn.py:23-24)
    n_fused = sum(isinstance(m, fused.FusedConvBNPLIF) for m in net.modules())
    assert n_fused == sum(eas.is_spiking_neuron(m) for m in net.modules()) == 34
    assert all(isinstance(m, fused.SeqToANNContainer) for m in net.dark5[1].m)


def test_training_path_channel_concat_keeps_channels_last_views():
    """fused._cat_channels == torch.cat(dim=-3) in value; on [T, B, C, H, W] views of channels-last buffers (the training
    path with cuDNN NHWC kernels) the result is again such a view -- no NCHW round trip between concat and the next conv."""
    import torch
    from eas_snn_b200 import fused
    T, B, H, W = 3, 2, 5, 7
    parts_cl = [torch.randn(T * B, c, H, W).contiguous(memory_format=torch.channels_last).view(T, B, c, H, W) for c in (4, 8)]
    out = fused._cat_channels(parts_cl)
    assert torch.equal(out, torch.cat(parts_cl, dim=-3))
    assert not out.is_contiguous() and out.permute(0, 1, 3, 4, 2).is_contiguous()
    assert out.flatten(0, 1).is_contiguous(memory_format=torch.channels_last)      # what the next conv sees
    parts = [p.contiguous() for p in parts_cl]
    out2 = fused._cat_channels(parts)
    assert out2.is_contiguous() and torch.equal(out2, out)
    mixed = fused._cat_channels([parts_cl[0], parts[1]])
    assert torch.equal(mixed, out)
