"""GPU parity: the fused sampler kernels (through the C ABI / module surface) against the golden
vectors from the reference and against the dense oracle on seeded inputs.

Tolerance (hard part 3): |a-b| <= 1e-5*max(1,|b|) on outputs, compared where the
spike histories agree; spike-history mismatch budget 1e-4 (near-threshold float ties)."""
import numpy as np
import pytest
import torch

import eas_snn_b200 as eas
from eas_snn_b200 import synth
from oracle import sampler as osamp
from helpers import load_golden, sampler_case, sampler_kwargs, close_report

pytestmark = pytest.mark.gpu

NAMES = list(load_golden("sampler")["names"])


def _compare(out_gpu, out_ref, budget=1e-4):
    a, b = out_gpu.detach().cpu(), out_ref.detach().cpu()
    err = (a - b).abs()
    tol = 1e-5 * torch.clamp(b.abs(), min=1.0)
    bad = err > tol
    frac = bad.float().mean().item()
    msg = "max|d|=%.3e bad=%d/%d (%.2e)" % (err.max().item(), int(bad.sum()), bad.numel(), frac)
    if bad.any():
        idx = torch.nonzero(bad)[:6].tolist()
        msg += " first %s got %s want %s" % (idx, [a[tuple(i)].item() for i in idx], [b[tuple(i)].item() for i in idx])
    return frac <= budget, frac, msg


# "tensor" = row-folded tcgen05 kernel (inputs exact in one fp16 plane: event counts), "tensor_split" = the
# tcgen05 kernel with hi + lo input planes (real-valued inputs), "fp32" = FP32-pipe kernel
ALGOS = ["fp32", "tensor", "tensor_split"]


def _tc_capable(cfg_or_kw, W):
    k = cfg_or_kw.get("ksize", cfg_or_kw.get("kernel_size"))
    return k == 5 and cfg_or_kw["depth"] == 2 and W % 4 == 0


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("name", NAMES)
def test_forward_golden(cuda, name, algo):
    z = load_golden("sampler")
    cfg, params, grads, x, y = sampler_case(z, name)
    if algo != "fp32" and not _tc_capable(cfg, cfg["W"]):
        pytest.skip("the tensor-core kernel covers depth 2 / k 5 / W % 4 == 0")
    m = eas.AdaptiveRSNNEmbedding(**sampler_kwargs(cfg)).to(cuda)
    m.algo = algo
    m.load_state_dict(params)
    with torch.no_grad():
        out = m(x.to(cuda))
    assert out.shape == y.shape and out.dtype == torch.float32
    ok, frac, msg = _compare(out, y, budget=2e-3 if y.numel() < 10000 else 1e-3)
    assert ok, name + ": " + msg
    # the int32 front door gives the same result as the fp32 one
    with torch.no_grad():
        out_i = m(x.to(cuda).int())
    assert torch.equal(out, out_i)


@pytest.mark.parametrize("algo", ALGOS)
def test_forward_vs_oracle_gen1_shape(cuda, algo):
    """BASELINE config-1 shape: B=2, Tm=4, 240x304, published flags, Poisson counts."""
    torch.manual_seed(80)
    kw = dict(kernel_size=5, in_channel=2, out_channel=2, readout="sum", split=False, write_zero=True, abs=False,
              depth=2, nb_steps=4, vreset=0, thresh=1, embedding="arsnn", Ts=1, spike_attach=True)
    ref = osamp.OracleSampler(**kw)
    g = torch.Generator().manual_seed(99)
    x = torch.poisson(torch.full((2, 4, 2, 240, 304), 1.2), generator=g)
    with torch.no_grad():
        want, st = osamp.sampler_forward(x, *_wb(ref), Ts=1, thresh=1, vreset=0, readout="sum",
                                         spike_attach=True, write_zero=True, return_state=True)
    fired = (st["seg"] > 0).float().mean().item()
    assert 0.05 < fired < 0.95, fired
    m = eas.AdaptiveRSNNEmbedding(**kw).to(cuda)
    m.algo = algo
    m.load_state_dict(ref.state_dict())
    with torch.no_grad():
        out = m(x.to(cuda))
    ok, frac, msg = _compare(out, want, budget=1e-4)
    print(algo, msg)
    assert ok, msg


def _wb(ref):
    iw, ib = ref._wb(ref.input_conv)
    gw, gb = ref._wb(ref.gate_conv)
    return iw, ib, gw, gb


@pytest.mark.parametrize("algo", ALGOS)
def test_6d_input_and_batch_independence(cuda, algo):
    torch.manual_seed(1)
    m = eas.AdaptiveRSNNEmbedding(kernel_size=5, depth=2, nb_steps=4, thresh=1, vreset=0, Ts=1,
                                  write_zero=True, spike_attach=True).to(cuda)
    m.algo = algo
    x = torch.poisson(torch.full((3, 2, 4, 2, 32, 72), 1.3)).to(cuda)       # [B, Tl, Tm, 2, H, W]
    with torch.no_grad():
        full = m(x)
        assert full.shape == (1, 6, 2, 32, 72)
        one = m(x[1:2, 1])                                                   # window (b=1, l=1) alone
    assert torch.equal(full[:, 3], one[:, 0])


def test_forward_events_equals_bin_then_sample(cuda):
    H, W = 64, 80
    arrs = synth.make_batch(9, 3, H, W, 2e4, 6e4)
    d = [torch.from_numpy(a).to(cuda) for a in arrs]
    torch.manual_seed(2)
    m = eas.AdaptiveRSNNEmbedding(kernel_size=5, depth=2, nb_steps=4, thresh=1, vreset=0, Ts=1,
                                  write_zero=True, spike_attach=True).to(cuda)
    with torch.no_grad():
        a = m.forward_events(*d, H, W)
        hist = eas.bin_events(*d, H, W, 4)
        b = m(hist.float())
    assert a.shape == (1, 3, 2, H, W)
    assert torch.equal(a, b)
    assert (a != 0).float().mean().item() > 0.01


@pytest.mark.parametrize("shape", [(1, 4, 8, 12), (3, 4, 61, 100), (2, 5, 130, 236), (5, 3, 33, 480), (64, 2, 24, 32)])
@pytest.mark.parametrize("flags", [dict(readout="sum", Ts=1, vreset=0, spike_attach=True, write_zero=True, abs=False),
                                   dict(readout="avg", Ts=2, vreset=None, spike_attach=False, write_zero=False, abs=True),
                                   dict(readout="last", Ts=3, vreset=0.25, spike_attach=True, write_zero=False, abs=False)])
def test_tensor_kernel_vs_oracle_shapes(cuda, shape, flags):
    """Tensor-core kernel on ragged strip / segment geometries (1..5 strips per row, segments that
    start mid-image, fewer tiles than ring slots) against the dense oracle, and against the FP32 kernel."""
    B, Tm, H, W = shape
    torch.manual_seed(7)
    kw = dict(kernel_size=5, in_channel=2, out_channel=2, split=False, depth=2, nb_steps=Tm, thresh=1,
              embedding="arsnn", **flags)
    ref = osamp.OracleSampler(**kw)
    g = torch.Generator().manual_seed(B * 1000 + H)
    x = torch.poisson(torch.full((B, Tm, 2, H, W), 1.0), generator=g)
    x[0, 0, 0, H // 2, W // 2] = 300.0            # a count that is not exact in one bf16 plane
    real = H % 2 == 1
    if real:                                      # real-valued micro-frames (resized inputs): hi + lo input planes
        x = x * (0.5 + torch.rand(x.shape, generator=g))
    with torch.no_grad():
        want = ref(x)
    m = eas.AdaptiveRSNNEmbedding(**kw).to(cuda)
    m.load_state_dict(ref.state_dict())
    outs = {}
    # real-valued inputs are not exact in one fp16 plane: "auto" must notice and recompute on the FP32 kernel
    for algo in (["fp32", "auto", "tensor_split"] if real else ALGOS):
        m.algo = algo
        with torch.no_grad():
            outs[algo] = m(x.to(cuda))
        ok, frac, msg = _compare(outs[algo], want, budget=2e-3 if want.numel() < 20000 else 2e-4)
        print(algo, shape, msg)
        assert ok, algo + ": " + msg
    if real:
        assert torch.equal(outs["auto"], outs["fp32"])
    else:
        ok, frac, msg = _compare(outs["tensor"], outs["fp32"].cpu(), budget=2e-3 if want.numel() < 20000 else 2e-4)
        assert ok, "tensor vs fp32: " + msg
    ok, frac, msg = _compare(outs["tensor_split"], outs["fp32"].cpu(), budget=2e-3 if want.numel() < 20000 else 2e-4)
    assert ok, "tensor_split vs fp32: " + msg
    assert (want != 0).float().mean().item() > 0.01


def test_tensor_kernel_saves_same_sequences_for_backward(cuda):
    """Training forward: v_seq / gate_seq written by the tensor-core kernel feed the same backward."""
    torch.manual_seed(3)
    kw = dict(kernel_size=5, depth=2, nb_steps=4, thresh=1, vreset=0, Ts=1, write_zero=True, spike_attach=True)
    m = eas.AdaptiveRSNNEmbedding(**kw).to(cuda)
    x = torch.poisson(torch.full((2, 4, 2, 48, 64), 1.1)).to(cuda)
    grads = {}
    for algo in ALGOS:
        m.algo = algo
        m.zero_grad()
        out = m(x)
        (out * torch.linspace(0.5, 1.5, out.numel(), device=cuda).view_as(out)).sum().backward()
        grads[algo] = [p.grad.clone() for p in m.parameters()]
    for algo in ALGOS[1:]:
        for a, b in zip(grads[algo], grads["fp32"]):
            scale = b.abs().max().item() + 1e-12
            assert (a - b).abs().max().item() <= 2e-3 * scale, (algo, (a - b).abs().max().item(), scale)


def test_auto_falls_back_beyond_fp16_range(cuda):
    """The tensor-core kernels hold operands as fp16 planes; a count that is not exact in fp16 (row-folded kernel) or a
    magnitude >= 65504 raises the device flag and algo="auto" recomputes on the FP32-pipe kernel: bit-identical to
    algo="fp32"."""
    torch.manual_seed(5)
    kw = dict(kernel_size=5, depth=2, nb_steps=4, thresh=1, vreset=0, Ts=1, write_zero=True, spike_attach=True)
    m = eas.AdaptiveRSNNEmbedding(**kw).to(cuda)
    x = torch.poisson(torch.full((2, 4, 2, 40, 64), 1.0)).to(cuda)
    outs = {}
    for big in (0.0, 2049.0, 0.3, 70000.0):
        if big:
            x[1, 2, 0, 17, 33] = big
        for algo in ("auto", "fp32"):
            m.algo = algo
            with torch.no_grad():
                outs[algo] = m(x)
        if big:
            assert torch.equal(outs["auto"], outs["fp32"])
        else:   # in range: auto is the tensor-core kernel (fp32-equivalent, not bit-identical)
            ok, frac, msg = _compare(outs["auto"], outs["fp32"].cpu(), budget=2e-3)
            assert ok, msg


def test_back_to_back_forwards_are_deterministic_and_do_not_stall(cuda):
    """Stress: hundreds of forwards queued without a host sync in between (the way bench.py and a serving loop drive the
    kernel).  The warp-specialised kernels hand tiles over through mbarriers; an ordering bug there shows up as a
    watchdog trap (launch failure) or as run-to-run differences only under this kind of back-to-back load -- a
    trailing-tile parity aliasing in the row-folded kernel was found exactly this way."""
    torch.manual_seed(11)
    kw = dict(kernel_size=5, depth=2, nb_steps=4, thresh=1, vreset=0, Ts=1, write_zero=True, spike_attach=True)
    m = eas.AdaptiveRSNNEmbedding(**kw).to(cuda)
    for shape in ((64, 4, 240, 304), (7, 4, 61, 100), (16, 4, 360, 640)):
        x = torch.poisson(torch.full(shape[:2] + (2,) + shape[2:], 0.8)).to(cuda)
        for algo in ("tensor", "tensor_split"):
            m.algo = algo
            with torch.no_grad():
                first = m(x).clone()
                for _ in range(150):
                    out = m(x)
            torch.cuda.synchronize()
            assert torch.equal(out, first), (shape, algo)


def test_1mpx_rvt_layout_event_sum_and_sampler(cuda):
    """BASELINE config 3: RVT-preprocessed uint8 [n, 20, 360, 640] -> 'event_sum' counts (rvt_gen4.py:120-122,
    bit-exact vs numpy) -> Tm consecutive slices as the sampler's steps at 360x640 (6 strips of the
    tensor-core kernel) vs the oracle."""
    g = np.random.default_rng(3)
    Tm = 4
    rep = (g.random((Tm, 20, 360, 640)) < 0.04).astype(np.uint8) * g.integers(1, 9, (Tm, 20, 360, 640), dtype=np.uint8)
    rep[1, 3, 100, 200] = 255
    want_sum = rep.reshape(Tm, 2, -1, 360, 640).sum(axis=2)              # the reference's two lines
    got_sum = eas.rvt_event_sum(torch.from_numpy(rep).to(cuda), 10)
    assert np.array_equal(got_sum.cpu().numpy(), want_sum.astype(np.float32))
    odd = eas.rvt_event_sum(torch.from_numpy(rep[:, :, :37, :53].copy()).to(cuda), 10)    # ragged plane size
    assert np.array_equal(odd.cpu().numpy(), rep[:, :, :37, :53].reshape(Tm, 2, -1, 37, 53).sum(axis=2).astype(np.float32))
    torch.manual_seed(80)
    kw = dict(kernel_size=5, in_channel=2, out_channel=2, readout="sum", split=False, write_zero=True, abs=False,
              depth=2, nb_steps=Tm, vreset=0, thresh=1, embedding="arsnn", Ts=1, spike_attach=True)
    ref = osamp.OracleSampler(**kw)
    x = got_sum.unsqueeze(0)                                              # [B=1, Tm, 2, 360, 640]
    with torch.no_grad():
        want = ref(x.cpu())
    m = eas.AdaptiveRSNNEmbedding(**kw).to(cuda)
    m.load_state_dict(ref.state_dict())
    for algo in ALGOS:
        m.algo = algo
        with torch.no_grad():
            out = m(x)
        ok, frac, msg = _compare(out, want, budget=1e-4)
        print(algo, msg)
        assert ok, algo + ": " + msg
    assert (want != 0).float().mean().item() > 0.01


def test_spike_count_embedding_golden_and_events(cuda):
    """(f-4) SpikeCountEmbedding: the reference class's outputs (tests/golden/count.npz), int32 histograms, and
    raw events -> count frames = total events per pixel and polarity inside the Tm micro-bins (bit-exact)."""
    from helpers import load_golden
    from oracle import binning as ob
    z = load_golden("count")
    m = eas.SpikeCountEmbedding(4)
    assert not list(m.parameters())
    for k in ("5", "6", "4"):
        x = torch.from_numpy(z["x" + k].astype(np.float32)).to(cuda)
        assert torch.equal(m(x).cpu(), torch.from_numpy(z["y" + k])), k
    xi = torch.from_numpy(z["x5"].astype(np.int32)).to(cuda)
    assert torch.equal(m(xi).cpu(), torch.from_numpy(z["y5"]))
    x, y, t, p, off = synth.make_batch(4, 5, 240, 304, 2e4, 6e4)
    got = m.forward_events(*(torch.from_numpy(a).to(cuda) for a in (x, y, t, p, off)), 240, 304)
    want = ob.micro_sum_batch(x, y, t, p, off, 240, 304, 4).sum(axis=1)
    assert got.shape == (5, 2, 240, 304) and np.array_equal(got.cpu().numpy(), want.astype(np.float32))


def _ablation_case(z, name):
    import ast
    cfg = ast.literal_eval(str(z[name + "/cfg"][0]))
    sd = {k[len(name) + 4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(name + "/sd/")}
    return cfg, sd, torch.from_numpy(z[name + "/x"].astype(np.float32)), torch.from_numpy(z[name + "/y"])


def test_ablation_embeddings_golden(cuda):
    """(f-4) ``LIFEmbedding`` ('snn') and ``SpikingEmbedding`` ('rsnn') -- strict subsets of the sampler's arithmetic,
    run on the sampler kernels -- against the reference classes' outputs; their state dicts load key for key
    (``embedding_conv.layer.0.weight``, ``cell.decay``, ``input_conv.layer.2.bias``, ``gate_conv.0.weight``)."""
    z = load_golden("ablations")
    for name in [str(n) for n in z["names"]]:
        cfg, sd, x, want = _ablation_case(z, name)
        vreset = None if cfg["vreset"] < -1e29 else cfg["vreset"]
        kw = dict(nb_steps=4, vreset=vreset, thresh=1, decay=torch.nn.Parameter(torch.tensor(0.0)), Ts=1)
        if cfg["kind"] == "lif":
            m = eas.LIFEmbedding(kernel_size=cfg["ksize"], readout=cfg["readout"], depth=cfg["depth"], **kw)
        else:
            m = eas.SpikingEmbedding(kernel_size=cfg["ksize"], readout=cfg["readout"], relu=cfg["relu"],
                                     depth=cfg["depth"], **kw)
        assert set(m.state_dict()) == set(sd), (name, set(m.state_dict()) ^ set(sd))
        m.load_state_dict(sd, strict=True)
        m = m.to(cuda).eval()
        for algo in (["fp32", "auto"] if cfg["ksize"] == 5 and cfg["depth"] == 2 else ["fp32"]):
            m.algo = algo
            with torch.no_grad():
                got = m(x.to(cuda))
            ok, frac, msg = _compare(got, want, budget=2e-3)
            print(name, algo, msg)
            assert got.shape == want.shape and ok, name + " " + algo + ": " + msg


def test_record_and_v_record_golden(cuda):
    """The analysis outputs ``forward(events, record=True)`` / ``v_record=True`` (embedding.py:198-199, 221-224; the
    Fig. 4 study of the README) against the reference's: the t_last history exactly, the sub-threshold potentials 1e-5."""
    z = load_golden("ablations")
    for name, Ts in (("record_ts1", 1), ("record_ts2", 2)):
        sd = {k[len(name) + 4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(name + "/sd/")}
        x = torch.from_numpy(z[name + "/x"].astype(np.float32)).to(cuda)
        m = eas.AdaptiveRSNNEmbedding(kernel_size=5, depth=2, nb_steps=5, thresh=1, vreset=0, Ts=Ts, write_zero=True,
                                      spike_attach=True).to(cuda)
        m.load_state_dict(sd, strict=True)
        m.algo = "fp32"
        y, rec = m(x, record=True)
        y2, vrec = m(x, v_record=True)
        want_rec = torch.from_numpy(z[name + "/record"].astype(np.int64))
        assert rec.shape == want_rec.shape and rec.dtype == torch.int64
        assert float((rec.cpu() != want_rec).float().mean()) <= 1e-4, name
        ok, frac, msg = _compare(y, torch.from_numpy(z[name + "/y"]), budget=2e-3)
        assert ok and torch.equal(y, y2), msg
        want_v = torch.from_numpy(z[name + "/v_record"])
        assert vrec.shape == want_v.shape, (vrec.shape, want_v.shape)      # same spikes -> same number of entries
        assert torch.allclose(vrec.cpu(), want_v, rtol=1e-5, atol=1e-5)


def test_voxel_grid_golden_and_oracle(cuda):
    """(f-4) ``eas.voxel_grid`` (to_voxel_grid_numpy, event_reps.py:30-89): the reference's outputs on its own bool-polarity
    dtype (every event +1) and on a signed dtype (+1 / -1); several windows in one call vs the oracle; a degenerate window."""
    from oracle import reps
    z = load_golden("voxel")
    i = 0
    while "%d/cfg" % i in z.files:
        n, H, W, nb = (int(v) for v in z["%d/cfg" % i])
        x, y, t, p = synth.make_window(np.random.default_rng(500 + i), n, H, W)
        d = [torch.from_numpy(a).to(cuda) for a in (x, y, t, p, np.array([0, n], np.int64))]
        for tag, mode in (("bool", "reference"), ("int8", "signed")):
            if "%d/%s" % (i, tag) in z.files:
                got = eas.voxel_grid(*d, H, W, nb, polarity=mode)
                want = torch.from_numpy(z["%d/%s" % (i, tag)]).float().unsqueeze(0)
                assert got.shape == want.shape and torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-5), (i, tag)
            elif n == 1:
                assert float(eas.voxel_grid(*d, H, W, nb, polarity=mode).abs().sum()) == 0.0   # t_last == t_first
        i += 1
    arrs = synth.make_batch(31, 5, 120, 152, 2e4, 9e4)
    got = eas.voxel_grid(*(torch.from_numpy(a).to(cuda) for a in arrs), 120, 152, 6, polarity="signed").cpu()
    off = arrs[4]
    for b in range(5):
        sl = slice(off[b], off[b + 1])
        want = reps.to_voxel_grid(arrs[0][sl], arrs[1][sl], arrs[2][sl], arrs[3][sl], 120, 152, 6, p_is_bool=False)
        assert torch.allclose(got[b], torch.from_numpy(want).float(), rtol=1e-5, atol=2e-5), b


# ---- the compact byte histogram as the sampler's input (in_dtype EAS_U8) ------------------------------------------
@pytest.mark.parametrize("flags", [dict(readout="sum", Ts=1, vreset=0, spike_attach=True, write_zero=True, abs=False),
                                   dict(readout="avg", Ts=2, vreset=None, spike_attach=False, write_zero=False, abs=True)])
def test_compact_histogram_input_is_bit_identical_to_dense(cuda, flags):
    """Same counts, same fp16 operands, same kernel: the byte histogram (incl. saturated bins served from its list,
    and a count beyond fp16's exact range that sends both down the FP32 fall-back) gives the dense input's frames
    bit for bit -- on ragged geometries and at the Gen1 shape."""
    torch.manual_seed(21)
    for (B, Tm, H, W), hot in (((3, 4, 61, 100), (255, 300)), ((8, 4, 240, 304), (254, 255, 1999)), ((2, 3, 33, 480), (5000,)),
                               ((2, 4, 40, 64), ())):
        m = eas.AdaptiveRSNNEmbedding(kernel_size=5, depth=2, nb_steps=Tm, thresh=1, **flags).to(cuda)
        rng = np.random.default_rng(B * 7 + H)
        n = int(B * H * W * 1.5)
        sizes = rng.multinomial(n, np.ones(B) / B)
        parts = [synth.make_window(rng, int(k), H, W) for k in sizes]
        x, y, t, p = (np.concatenate([q[i] for q in parts]) for i in range(4))
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        for j, n_hot in enumerate(hot):                # hot pixels in the first micro-bin of window 0 and the last of B-1
            s, e = (int(off[0]), int(off[1])) if j % 2 == 0 else (int(off[B - 1]), int(off[B]))
            sel = np.arange(s, s + n_hot) if j % 2 == 0 else np.arange(e - n_hot - 1, e - 1)
            x[sel], y[sel], p[sel] = 7 + 4 * j, 5 + j, j & 1
        d = [torch.from_numpy(a).to(cuda) for a in (x, y, t, p, off)]
        ch = eas.bin_events(*d, H, W, Tm, dtype=torch.uint8)
        dense = eas.bin_events(*d, H, W, Tm, dtype=torch.float32)
        assert torch.equal(ch.dense(torch.float32), dense)
        if hot:
            assert int(ch.tail[0]) >= 1 and int(dense.max()) >= 255
        for algo in ("auto", "fp32"):
            m.algo = algo
            with torch.no_grad():
                a, b = m(ch), m(dense)
            assert torch.equal(a, b), ((B, Tm, H, W), algo, float((a - b).abs().max()))
        m.algo = "auto"
        with torch.no_grad():
            b = m(dense)
        with torch.no_grad():
            c = m.forward_events(*d, H, W, hist_dtype="compact")
            e_ = m.forward_events(*d, H, W, hist_dtype="dense")
            f_ = m.forward_events(*d, H, W)
        assert torch.equal(c, b) and torch.equal(e_, b) and torch.equal(f_, b)


def test_compact_histogram_training_and_unsupported_shapes_take_the_dense_path(cuda):
    torch.manual_seed(22)
    H, W = 30, 50                                       # W % 4 != 0: FP32-pipe sampler; the byte form is still accepted
    arrs = synth.make_batch(9, 2, H, W, 3e3, 6e3)
    d = [torch.from_numpy(a).to(cuda) for a in arrs]
    m = eas.AdaptiveRSNNEmbedding(kernel_size=5, depth=2, nb_steps=4, thresh=1, vreset=0, Ts=1).to(cuda)
    ch = eas.bin_events(*d, H, W, 4, dtype=torch.uint8)
    with torch.no_grad():
        assert torch.equal(m(ch), m(ch.dense(torch.float32)))
        assert torch.equal(m.forward_events(*d, H, W), m(ch.dense(torch.float32)))
    with pytest.raises(ValueError):
        m.forward_events(*d, H, W, hist_dtype="compact")
    out = m(ch)                                         # grad mode: dense path, differentiable w.r.t. the parameters
    out.sum().backward()
    assert m.input_conv[0].weight.grad is not None and float(m.input_conv[0].weight.grad.abs().sum()) > 0


def test_tensor_kernel_randomised_geometries_vs_fp32_kernel(cuda):
    """Seeded sweep over what shapes the row-folded kernel's schedule: heights from one row to several row groups,
    1..6 strips per row, batches that give a CTA less than one segment or several, Tm = 1 (first-and-last step code),
    Tm up to the byte-packed limit, Ts up to 15, all three read-outs, byte / int32 / fp32 histograms -- against the
    FP32-pipe kernel (itself pinned to the oracle above) with the same 1e-5 bar."""
    rng = np.random.default_rng(2024)
    n_cases = 0
    for trial in range(36):
        H = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 9, 13, 31, 64, 97]))
        W = 4 * int(rng.choice([1, 2, 3, 7, 19, 29, 30, 58, 59, 77, 120, 160]))
        B = int(rng.choice([1, 2, 3, 5, 17, 40]))
        if B * H * W > 1_500_000:
            B = 2
        Tm = int(rng.choice([1, 2, 3, 4, 6, 14]))
        Ts = int(rng.choice([1, 1, 2, 3, 15]))
        flags = dict(readout=str(rng.choice(["sum", "avg", "last"])), Ts=Ts,
                     vreset=[0, None, 0.25][int(rng.integers(3))], spike_attach=bool(rng.integers(2)),
                     write_zero=bool(rng.integers(2)), abs=bool(rng.integers(2)))
        torch.manual_seed(100 + trial)
        m = eas.AdaptiveRSNNEmbedding(kernel_size=5, depth=2, nb_steps=Tm, thresh=1, **flags).to(cuda)
        g = torch.Generator().manual_seed(trial)
        x = torch.poisson(torch.full((B, Tm, 2, H, W), float(rng.choice([0.3, 1.0, 2.5]))), generator=g).to(cuda)
        kind = trial % 3
        inp = x if kind == 0 else x.to(torch.int32)
        m.algo = "fp32"
        with torch.no_grad():
            want = m(inp)
            m.algo = "tensor"
            got = m(inp)
            if kind == 2 and eas.binning.compact_fits(H, W):   # the same counts as a compact byte histogram
                ch = eas.CompactHist.empty((B, Tm, 2, H, W), cuda)
                ch.buf.zero_()
                ch.counts.copy_(x.to(torch.uint8))
                assert torch.equal(m(ch), got), (trial, "compact")
        ok, frac, msg = _compare(got, want.cpu(), budget=2e-3 if want.numel() < 20000 else 3e-4)
        assert ok, "trial %d %s %s: %s" % (trial, (B, Tm, H, W), flags, msg)
        n_cases += 1
    assert n_cases == 36
