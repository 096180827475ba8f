"""GPU parity: eas_bin_events (through the C ABI) is bit-exact against the golden vectors, against
the numpy oracle on seeded synthetic windows, and satisfies size-independent properties at the
BASELINE sizes."""
import numpy as np
import pytest
import torch

import eas_snn_b200 as eas
from eas_snn_b200 import synth
from oracle import binning as ob
from helpers import load_golden

pytestmark = pytest.mark.gpu


def _dev(arrs, cuda):
    x, y, t, p, off = arrs
    return (torch.from_numpy(x).to(cuda), torch.from_numpy(y).to(cuda), torch.from_numpy(t).to(cuda),
            torch.from_numpy(p).to(cuda), torch.from_numpy(off).to(cuda))


@pytest.mark.parametrize("strategy", ["reds", "tiles", "tiles_pair", "tiles_planes", "auto"])
def test_golden_cases(cuda, strategy):
    z = load_golden("binning")
    for name in z["names"]:
        H, W, Tm = (int(v) for v in z[f"{name}/dims"])
        x, y, t, p = (z[f"{name}/{k}"] for k in "xytp")
        off = np.array([0, len(x)], np.int64)
        got = eas.bin_events(*_dev((x, y, t, p, off), cuda), H, W, Tm, strategy=strategy)
        assert got.dtype == torch.int32 and got.shape == (1, Tm, 2, H, W)
        want = z[f"{name}/hist"]
        g = got[0].cpu().numpy()
        assert np.array_equal(g, want), "%s/%s: %d bins differ, sum got %d want %d" % (
            name, strategy, int((g != want).sum()), int(g.sum()), int(want.sum()))


@pytest.mark.parametrize("strategy", ["reds", "tiles", "tiles_pair", "tiles_planes"])
def test_ragged_batch_vs_oracle(cuda, strategy):
    """Ragged batch incl. empty windows, single events and a window at the 16 B alignment edge."""
    rng = np.random.default_rng(11)
    H, W, Tm = 60, 76, 4
    sizes = [0, 1, 7, 4093, 0, 12000, 3, 25001, 0]
    parts = [synth.make_window(rng, n, H, W) for n in sizes]
    x, y, t, p = (np.concatenate([q[i] for q in parts]) for i in range(4))
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    want = ob.micro_sum_batch(x, y, t, p, off, H, W, Tm).astype(np.int32)
    got = eas.bin_events(*_dev((x, y, t, p, off), cuda), H, W, Tm, strategy=strategy).cpu().numpy()
    assert np.array_equal(got, want), "bins differ: %d" % int((got != want).sum())


def test_gen1_batch_vs_oracle_and_properties(cuda):
    """BASELINE config-2 shape: 16 Gen1-rate windows at 240x304, Tm=4 (oracle finishes in seconds)."""
    H, W = synth.GEN1
    arrs = synth.gen1_batch(16)
    x, y, t, p, off = arrs
    d = _dev(arrs, cuda)
    a = eas.bin_events(*d, H, W, 4, strategy="tiles")
    b = eas.bin_events(*d, H, W, 4, strategy="reds")
    assert torch.equal(a, b)
    want = ob.micro_sum_batch(x, y, t, p, off, H, W, 4).astype(np.int32)
    assert np.array_equal(a.cpu().numpy(), want)
    # conservation: per window, counted + tail-dropped == events
    per_win = a.sum(dim=(1, 2, 3, 4)).cpu().numpy()
    n_win = np.diff(off)
    assert np.all(per_win <= n_win) and np.all(n_win - per_win < 200)
    # polarity split: channel 1 total == number of counted p != 0 events
    assert int(a[:, :, 1].sum()) <= int(p.sum())
    # idempotence / determinism
    assert torch.equal(a, eas.bin_events(*d, H, W, 4, strategy="tiles"))
    # fp32 counts (what the sampler consumes) are the same numbers
    for s in ("tiles", "tiles_pair", "tiles_planes", "reds"):
        f = eas.bin_events(*d, H, W, 4, strategy=s, dtype=torch.float32)
        assert f.dtype == torch.float32 and torch.equal(f, a.float())
        assert torch.equal(eas.bin_events(*d, H, W, 4, strategy=s), a)


def test_mpx_window_properties(cuda):
    """720x1280 (gen4): one 2e6-event window; checked by linearity (hist(A ++ B) == hist(A)+hist(B)
    when both halves are binned with the same boundaries) and against the oracle."""
    H, W = synth.MPX
    rng = np.random.default_rng(5)
    x, y, t, p = synth.make_window(rng, 2_000_000, H, W)
    off = np.array([0, len(x)], np.int64)
    got = eas.bin_events(*_dev((x, y, t, p, off), cuda), H, W, 4)
    want = ob.micro_sum(x, y, t, p, H, W, 4).astype(np.int32)
    assert np.array_equal(got[0].cpu().numpy(), want)
    assert int(got.sum()) == int(want.sum())


def test_out_of_range_events_are_ignored(cuda):
    x = np.array([0, 5, -1, 8, 3], np.int16)
    y = np.array([0, 5, 2, 2, 9], np.int16)
    t = np.array([0, 10, 20, 30, 41], np.int64)
    p = np.array([1, 0, 1, 0, 1], np.uint8)
    off = np.array([0, 5], np.int64)
    for s in ("reds", "tiles"):
        got = eas.bin_events(*_dev((x, y, t, p, off), cuda), 8, 8, 4, strategy=s)
        assert int(got.sum()) == 2 and int(got[0, 0, 1, 0, 0]) == 1 and int(got[0, 1, 0, 5, 5]) == 1


def test_letterbox_frames_match_the_reference_resize(cuda):
    """(f-3) eas_letterbox_bilinear vs the reference's get_random_data(random=False) outputs (golden sample + checksum)
    and the full oracle restatement: fp32 arithmetic against cv2's float64 -> 1e-6 relative."""
    from helpers import letterbox_cases, load_golden
    from oracle import letterbox as ol
    for (ih, iw, h, w, center, lb), fr, sample, chk in letterbox_cases(load_golden("letterbox")):
        want = ol.letterbox_frames(fr, (h, w), lb, center)
        for dt in (torch.float32, torch.int32):
            got = eas.letterbox_frames(torch.from_numpy(fr).to(cuda).to(dt), (h, w), letterbox=lb, center=center)
            assert got.shape == (3, 2, h, w) and got.dtype == torch.float32
            g = got.cpu().double().numpy()
            err = np.abs(g - want) / np.maximum(1.0, np.abs(want))
            assert err.max() <= 1e-6, ((ih, iw, h, w), err.max())
            assert np.abs(g[:, :, ::3, ::5] - sample).max() <= 1e-5
            assert abs(g.sum() - chk[0]) <= 1e-6 * chk[0]
    # 5-D histograms keep their leading dims; the canvas outside the pasted image is zero
    hist = torch.poisson(torch.full((2, 4, 2, 240, 304), 0.5)).to(cuda)
    out = eas.letterbox_frames(hist, (640, 640))
    assert out.shape == (2, 4, 2, 640, 640) and not out[..., 505:, :].any() and out[..., :505, :].any()


# ---- the compact byte histogram (EAS_U8): same information as the int32 histogram --------------------------------
def _compact_vs_dense(d, H, W, Tm, want_i32, strategy="auto"):
    ch = eas.bin_events(*d, H, W, Tm, dtype=torch.uint8, strategy=strategy)
    assert isinstance(ch, eas.CompactHist) and ch.shape == tuple(want_i32.shape)
    ch.check()
    want = torch.from_numpy(want_i32)
    assert torch.equal(ch.counts.cpu(), want.clamp(max=255).to(torch.uint8))            # bytes: min(count, 255)
    assert torch.equal(ch.dense(torch.int32).cpu(), want)                                 # exact again through the list
    assert torch.equal(ch.dense(torch.float32).cpu(), want.float())
    idx, cnt = ch.saturated()
    sat = (want.flatten() >= 255).nonzero().flatten()
    order = torch.argsort(idx.cpu())
    assert torch.equal(idx.cpu()[order], sat) and torch.equal(cnt.cpu()[order].long(), want.flatten()[sat].long())
    return ch


def test_compact_histogram_golden_cases(cuda):
    """Every reference golden (incl. the pixel whose count passes 65535 -- several 16-bit chunks -- and tw == 0)."""
    z = load_golden("binning")
    n_sat = 0
    for name in z["names"]:
        H, W, Tm = (int(v) for v in z[f"{name}/dims"])
        x, y, t, p = (z[f"{name}/{k}"] for k in "xytp")
        off = np.array([0, len(x)], np.int64)
        want = z[f"{name}/hist"][None].astype(np.int32)
        if not eas.binning.compact_fits(H, W):
            with pytest.raises(RuntimeError):
                eas.bin_events(*_dev((x, y, t, p, off), cuda), H, W, Tm, dtype=torch.uint8)
            continue
        for strategy in ("tiles_planes", "tiles_pair"):
            ch = _compact_vs_dense(_dev((x, y, t, p, off), cuda), H, W, Tm, want, strategy)
        n_sat += int(ch.tail[0])
    assert n_sat > 0, "no golden exercised the saturation list"


def test_compact_histogram_hot_pixels_and_gen1_batch(cuda):
    """Gen1 batch with planted hot pixels at 254 / 255 / 256 / 3000 events per micro-bin, SoA and .dat front doors."""
    H, W = synth.GEN1
    x, y, t, p, off = synth.gen1_batch(8)
    rng = np.random.default_rng(5)
    xs, ys, ts, ps, sizes = [], [], [], [], []
    for b in range(8):
        s, e = int(off[b]), int(off[b + 1])
        xb, yb, tb, pb = x[s:e].copy(), y[s:e].copy(), t[s:e].copy(), p[s:e].copy()
        first_bin = np.nonzero(tb < tb[0] + (tb[-1] - tb[0]) // 4)[0]
        for k, n_hot in enumerate((254, 255, 256, 3000)):
            seg = len(first_bin) // 4
            sel = first_bin[k * seg:k * seg + min(n_hot, seg)]        # move these events onto one pixel / polarity
            xb[sel], yb[sel], pb[sel] = 10 + k, 20 + b, (k + b) & 1
        xs.append(xb), ys.append(yb), ts.append(tb), ps.append(pb), sizes.append(e - s)
    x, y, t, p = (np.concatenate(v) for v in (xs, ys, ts, ps))
    want = ob.micro_sum_batch(x, y, t, p, off, H, W, 4).astype(np.int32)
    assert (want >= 255).sum() >= 8 * 3
    d = _dev((x, y, t, p, off), cuda)
    for strategy in ("tiles_planes", "tiles_pair", "auto"):
        _compact_vs_dense(d, H, W, 4, want, strategy)
    rec = torch.from_numpy(eas.pack_records(x, y, t, p)).to(cuda)
    rng_ = torch.from_numpy(np.stack([off[:-1], off[1:]], 1)).to(cuda)
    ch = eas.bin_dat(rec, rng_, H, W, 4, dtype=torch.uint8)
    assert torch.equal(ch.dense().cpu(), torch.from_numpy(want))
    buf = eas.CompactHist.empty((8, 4, 2, H, W), cuda)            # caller-owned buffer, reused
    for _ in range(2):
        assert eas.bin_events(*d, H, W, 4, out=buf) is buf
        assert torch.equal(buf.dense().cpu(), torch.from_numpy(want))


def test_compact_histogram_reports_lost_counts(cuda):
    """More saturated bins than the list holds: the call says so (lazily through poll_compact, or .check())."""
    H, W, per = 64, 80, 256
    n = H * W * per
    x = np.tile(np.arange(W, dtype=np.int16), H * per)
    y = np.repeat(np.arange(H, dtype=np.int16), W * per)
    t = np.arange(n, dtype=np.int64)
    p = np.ones(n, np.uint8)
    order = np.random.default_rng(0).permutation(n)
    d = _dev((x[order], y[order], t, p, np.array([0, n], np.int64)), cuda)
    ch = eas.bin_events(*d, H, W, 1, dtype=torch.uint8)
    with pytest.raises(OverflowError):
        ch.check()
    with pytest.raises(OverflowError):
        eas.poll_compact(wait=True)
    eas.poll_compact(wait=True)                                     # reported once
    assert int(eas.bin_events(*d, H, W, 1).sum()) == n - 1          # the dense histogram is unaffected (last event: t == t0 + tw, dropped)
