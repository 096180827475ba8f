"""GPU parity: the tcgen05 conv -> folded BN -> PLIF kernel (through the C ABI) against fp32 PyTorch
convolution (pre-activation mode) and against the oracle's conv -> BN -> PLIF (spike mode), and the
whole fused spiking CSPDarknet against the golden vector produced by the reference model."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import eas_snn_b200 as eas
from eas_snn_b200 import fused
from oracle import backbone as ob, plif as op
from helpers import load_golden, close_report

pytestmark = pytest.mark.gpu


def _ref_conv(x_cl, w, stride):
    """x_cl [Tx,B,H,W,C] fp32 (exact values), w [Cout,Cin,k,k] fp32 -> [Tx,B,Ho,Wo,Cout] fp64-accumulated."""
    Tx, B, H, W, C = x_cl.shape
    k = w.shape[-1]
    y = F.conv2d(x_cl.flatten(0, 1).permute(0, 3, 1, 2).double(), w.double(), None, stride, (k - 1) // 2)
    return y.permute(0, 2, 3, 1).reshape(Tx, B, y.shape[2], y.shape[3], -1)


SHAPES = [
    # Tx, B, H,  W,  Cin, Cout, k, stride
    (1, 1, 8, 16, 64, 64, 1, 1),       # exactly one 128 x 64 x 64 tile
    (1, 2, 16, 24, 64, 64, 1, 1),      # several M tiles
    (3, 2, 16, 24, 128, 96, 1, 1),     # 2 K blocks, N tail
    (1, 1, 16, 16, 64, 64, 3, 1),      # 3x3 taps + zero padding by TMA
    (3, 2, 20, 28, 48, 80, 3, 1),      # Cin < 64 (channel OOB fill), ragged tiles
    (3, 2, 32, 40, 32, 64, 3, 2),      # stride 2 through tensor-map element strides
    (3, 4, 8, 10, 192, 192, 3, 1),     # small map: several images per tile
    (2, 3, 5, 7, 16, 24, 1, 1),        # tiny, everything ragged
    (3, 2, 16, 20, 8, 16, 3, 2),       # Cin = 8 (the golden backbone's first layers)
]


@pytest.mark.parametrize("shape", SHAPES)
def test_preact_matches_fp32_conv(cuda, shape):
    Tx, B, H, W, Cin, Cout, k, stride = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randint(0, 3, (Tx, B, H, W, Cin), generator=g).float()           # spikes / SEW sums
    w = torch.randn((Cout, Cin, k, k), generator=g) / (Cin * k * k) ** 0.5
    bias = torch.randn(Cout, generator=g)
    want = _ref_conv(x, w, stride) + bias.double()
    wp, un = fused.pack_weight(w.to(cuda), 2)
    got = fused.conv_bn_plif(x.to(cuda).half(), wp, bias.to(cuda), None, Tx, k, stride, out_mode=fused.OUT_PREACT,
                             w_unscale=un)
    assert got.shape == want.shape, (got.shape, want.shape)
    # tcgen05 accumulates in fp32 with truncation: the error grows ~1.6e-8 * K (K = taps * Cin); the fp16 hi/lo
    # weight split itself is exact to 2^-22.  (cuDNN's default TF32 path is ~1e-3 on the same data.)
    K = Cin * k * k
    ok, msg = close_report(got, want, rtol=1e-6, atol=2e-6 + 2.5e-8 * K)
    print("preact %s: %s" % (shape, msg))
    assert ok, "%s: %s" % (shape, msg)


@pytest.mark.parametrize("cfg", [
    # B, H,  W,  Cin, Cout, k, stride : single-accumulator layers big enough (>= 148 tiles) for 128-channel tiles
    (8, 64, 80, 64, 192, 3, 1),        # tap-reuse mode with two input planes, N tiles 128 + 64
    (8, 64, 80, 96, 96, 1, 1),         # 1x1, one 128-wide tile with 96 live channels
    (8, 64, 80, 64, 144, 3, 2),        # stride 2 (no tap reuse), N tail of 16
    (16, 32, 40, 192, 200, 3, 1),      # ragged Cout, 6 channel blocks in tap-reuse mode
])
def test_ann_layers_wide_tiles_and_split_input_reuse(cuda, cfg):
    """Real-valued input (two fp16 planes) -> SiLU planes / fp32 pre-activation on the wide-tile kernels."""
    B, H, W, Cin, Cout, k, stride = cfg
    g = torch.Generator().manual_seed(sum(cfg))
    x = torch.randn((1, B, H, W, Cin), generator=g)
    w = torch.randn((Cout, Cin, k, k), generator=g) / (Cin * k * k) ** 0.5
    bias = torch.randn(Cout, generator=g) * 0.2
    xc, bc = x.to(cuda), bias.to(cuda)
    wp, un = fused.pack_weight(w.to(cuda), 2)
    want = (F.conv2d(xc[0].permute(0, 3, 1, 2).double(), w.to(cuda).double(), bc.double(), stride, (k - 1) // 2)
            .permute(0, 2, 3, 1).unsqueeze(0))
    got = fused.conv_bn_plif(fused.split_f16(xc, 2), wp, bc, None, 1, k, stride, n_xsplit=2,
                             out_mode=fused.OUT_PREACT, w_unscale=un)
    K = Cin * k * k
    ok, msg = close_report(got, want, rtol=3e-6, atol=3e-6 + 2.5e-8 * K)
    print("wide tile %s: %s" % (cfg, msg))
    assert ok, msg
    planes = fused.conv_bn_plif(fused.split_f16(xc, 2), wp, bc, None, 1, k, stride, n_xsplit=2,
                                out_mode=fused.OUT_SILU2, w_unscale=un)
    ok, msg = close_report(planes.float().sum(0), F.silu(want).float(), rtol=1e-5, atol=1e-5)
    assert ok, msg
    # and as the broadcast first spiking conv: one accumulator, T = 3 LIF steps in the epilogue
    spikes = fused.conv_bn_plif(fused.split_f16(xc, 2), wp, bc, torch.zeros((), device=cuda), 3, k, stride,
                                n_xsplit=2, out_mode=fused.OUT_SPIKES, w_unscale=un)
    pre = want[0].float()
    v = torch.zeros_like(pre)
    for t in range(3):
        v = v * 0.5 + pre
        s_ = (v >= 1.0).float()
        v = v - s_
        mism = (spikes[t].float() != s_).float().mean().item()
        assert mism <= 1e-4, (t, mism)


def test_single_plane_is_fp16_accurate(cuda):
    g = torch.Generator().manual_seed(1)
    x = torch.randint(0, 2, (1, 1, 16, 16, 64), generator=g).float()
    w = torch.randn((64, 64, 3, 3), generator=g) / 24.0
    bias = torch.zeros(64)
    wp, un = fused.pack_weight(w.to(cuda), 1)
    got = fused.conv_bn_plif(x.to(cuda).half(), wp, bias.to(cuda), None, 1, 3, 1, out_mode=fused.OUT_PREACT, w_unscale=un)
    want = _ref_conv(x, (wp[0].float() * un.view(-1, 1, 1, 1)).permute(0, 3, 1, 2).cpu(), 1)   # the single fp16 plane
    ok, msg = close_report(got, want, rtol=1e-5, atol=1e-5)
    assert ok, msg
    ok, msg = close_report(got, _ref_conv(x, w, 1), rtol=2e-3, atol=2e-3)      # and ~2^-11 of the fp32 weights
    assert ok, msg


def test_real_valued_input_split(cuda):
    g = torch.Generator().manual_seed(2)
    x = torch.randn((1, 2, 12, 20, 8), generator=g)
    w = torch.randn((32, 8, 3, 3), generator=g) / 8.0
    bias = torch.randn(32, generator=g) * 0.1
    want = _ref_conv(x, w, 1) + bias.double()
    wp, un = fused.pack_weight(w.to(cuda), 2)
    got = fused.conv_bn_plif(fused.split_f16(x.to(cuda), 2), wp, bias.to(cuda), None,
                             1, 3, 1, n_xsplit=2, out_mode=fused.OUT_PREACT, w_unscale=un)
    ok, msg = close_report(got, want, rtol=3e-6, atol=3e-6)
    print("real-valued input:", msg)
    assert ok, msg
    # SiLU planes sum back to the fp32 activation
    planes = fused.conv_bn_plif(fused.split_f16(x.to(cuda), 2), wp, bias.to(cuda),
                                None, 1, 3, 1, n_xsplit=2, out_mode=fused.OUT_SILU2, w_unscale=un)
    assert planes.shape[0] == 2 and planes.dtype == torch.float16
    silu = F.silu(want).float()
    ok, msg = close_report(planes.float().sum(0), silu, rtol=1e-5, atol=1e-5)
    assert ok, msg


@pytest.mark.parametrize("cfg", [(64, 64, 1, 1, 3), (32, 64, 3, 2, 3), (96, 48, 3, 1, 3), (64, 64, 1, 1, 1),
                                 (64, 96, 3, 1, 5), (48, 40, 1, 1, 8), (96, 96, 3, 1, 4)])
def test_layer_vs_oracle_conv_bn_plif(cuda, cfg):
    """Teacher-forced single layer: identical spike input, conv -> BN(eval) -> PLIF oracle vs the fused kernel."""
    Cin, Cout, k, stride, T = cfg
    torch.manual_seed(Cin + Cout + k)
    ref = ob.SpikingBaseConv(Cin, Cout, k, stride, op.ATan(2.0))
    x = (torch.rand(T, 2, Cin, 24, 32) < 0.2).float()
    x = x + (torch.rand_like(x) < 0.05).float()                       # a few 2s (SEW sums)
    ob.calibrate_bn(ref, x)
    ref.act.w.data.fill_(0.3)
    with torch.no_grad():
        want = ref(x)
    rate = want.mean().item()
    assert 0.02 < rate < 0.9, rate
    m = fused.FusedConvBNPLIF(Cin, Cout, k, stride, eas.ATan(2.0)).to(cuda)
    m.load_state_dict(ref.state_dict())
    m.eval()
    with torch.no_grad():
        got_cl = m.run(x.to(cuda).permute(0, 1, 3, 4, 2).contiguous().half(), T)
        got = got_cl.permute(0, 1, 4, 2, 3).float().cpu()
        got_mod = m(x.to(cuda))                                        # drop-in [T,B,C,H,W] fp32 path
    mism = (got != want).float().mean().item()
    print("layer %s: spike mismatch %.3e (rate %.3f)" % (cfg, mism, rate))
    assert mism <= 1e-4, "spike mismatch %.3e (rate %.3f)" % (mism, rate)
    assert got_mod.dtype == torch.float32 and got_mod.shape == want.shape
    assert (got_mod.cpu() != want).float().mean().item() <= 1e-4


def test_output_into_concat_slice(cuda):
    g = torch.Generator().manual_seed(4)
    x = torch.randint(0, 2, (2, 1, 8, 16, 64), generator=g).float().to(cuda).half()
    w, un = fused.pack_weight((torch.randn((32, 64, 1, 1), generator=g) / 8).to(cuda), 2)
    bias = torch.randn(32, generator=g).to(cuda)
    pw = torch.tensor(0.0, device=cuda)
    cat = torch.full((2, 1, 8, 16, 96), 7.0, dtype=torch.float16, device=cuda)
    alone = fused.conv_bn_plif(x, w, bias, pw, 2, 1, 1, w_unscale=un)
    fused.conv_bn_plif(x, w, bias, pw, 2, 1, 1, out=cat[..., 32:64], w_unscale=un)
    assert torch.equal(cat[..., 32:64], alone)
    assert bool((cat[..., :32] == 7).all()) and bool((cat[..., 64:] == 7).all())
    # and a channel slice as INPUT
    wide = torch.zeros((2, 1, 8, 16, 160), dtype=torch.float16, device=cuda)
    wide[..., 64:128] = x
    assert torch.equal(fused.conv_bn_plif(wide[..., 64:128], w, bias, pw, 2, 1, 1, w_unscale=un), alone)


def test_residual_and_wide_tiles(cuda):
    """SEW shortcut fused in the epilogue; a layer large enough for the 128-wide N tile; K blocks of 32."""
    g = torch.Generator().manual_seed(9)
    for (Cin, Cout, k, H, W) in [(192, 192, 3, 16, 20), (96, 96, 3, 16, 24), (128, 256, 3, 8, 12)]:
        x = (torch.rand((3, 2, H, W, Cin), generator=g) < 0.3).float()
        w = torch.randn((Cout, Cin, k, k), generator=g) / (Cin * k * k) ** 0.5 * 2.0
        bias = torch.randn(Cout, generator=g) * 0.3 + 0.3
        pre = (_ref_conv(x, w, 1) + bias.double()).float()
        want = op.plif_forward(pre, torch.tensor(0.2), op.ATan(2.0), 1.0, None, False, False)
        res = torch.randint(0, 3, want.shape, generator=g).float()
        xg, (wp, un) = x.to(cuda).half(), fused.pack_weight(w.to(cuda), 2)
        pw = torch.tensor(0.2, device=cuda)
        got = fused.conv_bn_plif(xg, wp, bias.to(cuda), pw, 3, k, 1, w_unscale=un).float().cpu()
        got_r = fused.conv_bn_plif(xg, wp, bias.to(cuda), pw, 3, k, 1, residual=res.to(cuda).half(),
                                   w_unscale=un).float().cpu()
        mism = (got != want).float().mean().item()
        print("wide/residual (%d,%d,%d): mismatch %.3e rate %.3f" % (Cin, Cout, k, mism, want.mean().item()))
        assert 0.02 < want.mean().item() < 0.95
        assert mism <= 1e-4
        assert torch.equal(got_r, got + res)


def test_backbone_golden(cuda):
    """Whole spiking CSPDarknet (tiny width) vs the golden spikes of the REFERENCE model."""
    z = load_golden("backbone")
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    net = fused.SpikingCSPDarknet(0.33, 0.125, in_dim=2, spike_fn=eas.ATan(2.0), T=3).to(cuda)
    net.load_state_dict(sd, strict=True)
    net.eval()
    x = torch.from_numpy(z["x"]).to(cuda)
    outs = net(x[:1])                                      # Ts == 1 frame, broadcast inside
    outs_T = net(x)                                        # explicit T frames
    for k in ("dark3", "dark4", "dark5"):
        want = torch.from_numpy(z["out/" + k]).float()
        got = outs[k].float().cpu()
        mism = (got != want).float().mean().item()
        print("backbone %s: spike mismatch %.3e (rate %.3f)" % (k, mism, want.mean().item()))
        assert got.shape == want.shape
        assert mism <= 2e-3, "%s mismatch %.3e (rate %.3f)" % (k, mism, want.mean().item())
        assert torch.equal(outs_T[k].float().cpu(), got), k + ": broadcast path differs from explicit T frames"


@pytest.mark.parametrize("shape", [(5, 8, 10, 64), (3, 12, 20, 40), (2, 3, 4, 8)])
def test_spp_pools_match_torch(cuda, shape):
    """eas_spp_pool_fwd (5 / 9 / 13 max-pools into the concat slices) against torch.max_pool2d: exact."""
    import torch.nn.functional as F
    from eas_snn_b200 import _lib
    N, H, W, C = shape
    g = torch.Generator().manual_seed(N + H)
    x = (torch.rand((N, H, W, C), generator=g) < 0.08).half() + (torch.rand((N, H, W, C), generator=g) < 0.02).half()
    cat = torch.full((N, H, W, 4 * C + 8), 7.0, dtype=torch.float16, device=cuda)
    cat[..., :C] = x.to(cuda)
    rc = _lib.lib().eas_spp_pool_fwd(_lib.ptr(cat), N, H, W, C, 4 * C + 8, 5, 9, 13, _lib.stream_ptr())
    assert rc == 0
    x4 = x.permute(0, 3, 1, 2).float()
    for i, k in enumerate((5, 9, 13)):
        want = F.max_pool2d(x4, k, 1, k // 2).permute(0, 2, 3, 1)
        assert torch.equal(cat[..., (i + 1) * C:(i + 2) * C].float().cpu(), want), k
    assert torch.equal(cat[..., :C].cpu(), x) and bool((cat[..., 4 * C:] == 7).all())


def test_graphed_backbone_replays_the_eager_forward(cuda):
    """CUDA-graph capture of the fused spiking CSPDarknet: replay on new frames equals the eager forward."""
    torch.manual_seed(5)
    net = fused.SpikingCSPDarknet(0.33, 0.25, in_dim=2, T=3).to(cuda).eval()
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.bias.data.fill_(0.6)
    x0 = torch.rand(1, 2, 2, 64, 96, device=cuda) * 2
    g = fused.GraphedForward(net, x0)
    for seed in (1, 2):
        x = torch.rand(1, 2, 2, 64, 96, device=cuda, generator=torch.Generator(cuda).manual_seed(seed)) * 2
        want = {k: v.clone() for k, v in net(x).items()}
        got = g(x)
        for k in want:
            assert torch.equal(got[k], want[k]), k
        assert 0.01 < float(want["dark5"].float().mean()) < 0.9


def test_focus_im2col_and_stem_match_torch(cuda):
    """eas_focus_im2col (space-to-depth + 3x3 im2col + fp16 hi/lo split) against torch slicing + unfold, and the
    stem (Focus conv -> BN -> SiLU, network_blocks.py:191-213) through it against fp32 PyTorch."""
    from eas_snn_b200 import _lib
    g = torch.Generator().manual_seed(9)
    Tx, B, H, W = 1, 3, 20, 28
    fr = (torch.randn((Tx, B, 2, H, W), generator=g) * 3).to(cuda)
    cols = torch.empty((2, Tx, B, H // 2, W // 2, 80), dtype=torch.float16, device=cuda)
    rc = _lib.lib().eas_focus_im2col(_lib.ptr(fr), Tx * B, H, W, _lib.ptr(cols), cols.stride(0), _lib.stream_ptr())
    assert rc == 0
    x = fr[0]
    s2d = torch.cat((x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]), dim=1)   # [B, 8, H/2, W/2]
    unf = F.unfold(s2d, 3, padding=1).view(B, 8, 9, H // 2, W // 2).permute(0, 3, 4, 2, 1).reshape(B, H // 2, W // 2, 72)
    got = cols.float().sum(0)[0]
    assert torch.allclose(got[..., :72], unf, rtol=2.0 ** -21, atol=1e-7)
    assert not got[..., 72:].any()
    torch.manual_seed(3)
    stem = fused._Focus(2, 24, 3).to(cuda).eval()
    stem.conv.bn.running_mean.normal_(0, 0.2)
    stem.conv.bn.running_var.uniform_(0.5, 1.5)
    stem.conv.bn.weight.data.uniform_(0.8, 1.2)
    with torch.no_grad():
        want = stem(x.double()) if False else stem.double()(x.double()).float()
        stem.float()
        planes = stem.run(fr)
    gotp = planes.float().sum(0)[0].permute(0, 3, 1, 2)
    ok, msg = close_report(gotp, want, rtol=1e-5, atol=1e-5)
    print("stem via im2col:", msg)
    assert ok, msg


def test_convert_to_spiking_model_reproduces_the_reference_backbone_golden(cuda):
    """The drop-in seam end to end: a plain ANN CSPDarknet -> ``fused.convert_to_spiking`` -> the reference's converted
    weights (backbone.npz, strict) -> its own module-by-module forward on [T, B, C, H, W] (each fused layer runs the
    tensor-core kernel, Focus / pools / concatenations / SEW adds stay PyTorch) -> the reference's spikes."""
    from eas_snn_b200 import fused
    from helpers import AnnCSPDarknet, load_golden
    z = load_golden("backbone")
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    net = fused.convert_to_spiking(AnnCSPDarknet(0.33, 0.125, in_dim=2), eas.ATan(2.0))
    net.load_state_dict(sd, strict=True)
    for m in net.modules():          # init_yolo (event_yolox_base.py:179-183), applied AFTER the conversion like get_model
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eps, m.momentum = 1e-3, 0.03
    net = net.to(cuda).eval()
    x = torch.from_numpy(z["x"]).to(cuda)
    with torch.no_grad():
        outs = net(x)
    eas.reset_net(net)
    for k, v in outs.items():
        want = torch.from_numpy(z["out/" + k].astype(np.float32))
        mism = float((v.float().cpu() != want).float().mean())
        print("convert_to_spiking", k, "spike mismatch %.2e (rate %.3f)" % (mism, float(want.mean())))
        assert v.shape == want.shape and mism <= 2e-3, (k, mism)
