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


@pytest.mark.parametrize("strategy", ["reds", "tiles", "auto"])
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


@pytest.mark.parametrize("strategy", ["reds", "tiles"])
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
    for s in ("tiles", "reds"):
        f = eas.bin_events(*d, H, W, 4, strategy=s, dtype=torch.float32)
        assert f.dtype == torch.float32 and torch.equal(f, a.float())


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
