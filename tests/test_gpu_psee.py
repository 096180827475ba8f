"""GPU parity of the PSEE .dat path: window search and binning on raw 8-byte records
against the golden output of the reference's own loader (tests/golden/psee.npz) and against the oracle
on seeded recordings.  Integer work: bit-exact."""
import numpy as np
import pytest
import torch

import eas_snn_b200 as eas
from eas_snn_b200 import synth
from oracle import psee as opsee, binning as obin
from helpers import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(synth.DAT_CASES))
def test_windows_and_histograms_match_reference_loader(cuda, name):
    z = load_golden("psee")
    kw, window, num_slice, Tm = synth.DAT_CASES[name]
    x, y, t, p = synth.dat_stream(**kw)
    H, W = kw["H"], kw["W"]
    recd = eas.DatRecording.from_events(x, y, t, p, cuda, H, W)
    ts = synth.dat_label_times(name, t, window)
    ranges = recd.windows(ts, window, num_slice)
    r = ranges.cpu().numpy()
    count = r[:, 1] - r[:, 0]
    assert np.array_equal(count, z[f"{name}/count"])
    nz = count > 0
    assert np.array_equal(r[nz, 0], z[f"{name}/first"][nz])
    for strategy in ("tiles", "tiles_pair", "tiles_planes", "reds", "auto"):
        for dtype in (torch.int32, torch.float32):
            hist = recd.histograms(ranges, H, W, Tm, strategy=strategy, dtype=dtype)
            assert np.array_equal(hist.cpu().numpy().astype(np.int32), z[f"{name}/hist"]), (strategy, dtype)


def test_bin_dat_equals_bin_events_on_decoded_records(cuda):
    """Same windows through both front doors (packed records vs the decoded SoA arrays), Gen1 frame size,
    overlapping and empty ranges, an out-of-frame coordinate."""
    H, W, Tm = 240, 304, 4
    x, y, t, p, off = synth.make_batch(21, 5, H, W, 2e4, 9e4)
    x = x.copy()
    x[17] = 400                                             # outside the frame: dropped by both paths
    rec = torch.from_numpy(eas.pack_records(x, y, t - t.min(), p)).to(cuda)
    ranges = torch.tensor([[off[b], off[b + 1]] for b in range(5)] + [[100, 100], [off[1] - 500, off[1] + 700]],
                          dtype=torch.int64, device=cuda)
    got = eas.bin_dat(rec, ranges, H, W, Tm)
    d = [torch.from_numpy(a).to(cuda) for a in (x, y, t - t.min(), p)]
    for b in range(ranges.shape[0]):
        lo, hi = (int(v) for v in ranges[b])
        o = torch.tensor([0, hi - lo], dtype=torch.int64, device=cuda)
        want = eas.bin_events(d[0][lo:hi].clone(), d[1][lo:hi].clone(), d[2][lo:hi].clone(),
                              d[3][lo:hi].clone(), o, H, W, Tm)      # clone: the SoA door wants 16 B aligned arrays
        assert torch.equal(got[b], want[0]), b
    want0 = obin.micro_sum(x[off[0]:off[1]], y[off[0]:off[1]], (t - t.min())[off[0]:off[1]], p[off[0]:off[1]], H, W, Tm)
    keep = x[off[0]:off[1]] < W
    assert int(got[0].sum()) <= int(keep.sum())
    assert np.array_equal(got[0].cpu().numpy(), want0.astype(np.int32)) or x[17] >= W   # oracle has no bounds check


def test_large_recording_windows_vs_oracle(cuda):
    """1.2 M records (several bisection probes), 200 random label times incl. exact probe hits."""
    rng = np.random.default_rng(5)
    n = 1_200_000
    t = np.sort(rng.integers(0, 4_000_000, n))
    x, y, p = rng.integers(0, 304, n), rng.integers(0, 240, n), rng.integers(0, 2, n)
    rec = opsee.pack_records(x, y, t, p)
    ts = list(rng.integers(-100_000, 4_300_000, 180))
    lo, hi = 0, n
    while hi - lo > 100000:
        mid = (lo + hi) // 2
        ts += [int(t[mid]) + 50_000, int(t[mid]) + 50_001]
        lo = mid + 1
    ts = np.asarray(ts, dtype=np.int64)
    want = opsee.windows(rec, ts, (-50_000, 0), 2)
    recd = eas.DatRecording.from_events(x, y, t, p, cuda)
    got = recd.windows(ts, (-50_000, 0), 2).cpu().numpy()
    assert np.array_equal(got[:, 1] - got[:, 0], want[:, 1] - want[:, 0])
    nz = want[:, 1] > want[:, 0]
    assert np.array_equal(got[nz], want[nz])
    hist = recd.histograms(torch.from_numpy(want[:8]).to(cuda), 240, 304, 4)
    assert np.array_equal(hist.cpu().numpy(), opsee.micro_sum_windows(rec, want[:8], 240, 304, 4).astype(np.int32))


def test_forward_dat_equals_forward_events(cuda):
    H, W = 64, 80
    x, y, t, p, off = synth.make_batch(9, 3, H, W, 2e4, 6e4)
    torch.manual_seed(2)
    m = eas.AdaptiveRSNNEmbedding(kernel_size=5, depth=2, nb_steps=4, thresh=1, vreset=0, Ts=1,
                                  write_zero=True, spike_attach=True).to(cuda)
    rec = torch.from_numpy(eas.pack_records(x, y, t, p)).to(cuda)
    ranges = torch.tensor([[off[b], off[b + 1]] for b in range(3)], dtype=torch.int64, device=cuda)
    d = [torch.from_numpy(a).to(cuda) for a in (x, y, t, p, off)]
    from eas_snn_b200.psee import bin_dat
    ha = bin_dat(rec, ranges, H, W, 4, dtype=torch.float32)
    hb = eas.bin_events(*d, H, W, 4, dtype=torch.float32)
    assert torch.equal(ha, hb), "fp32 histograms differ: %d bins" % int((ha != hb).sum())
    with torch.no_grad():
        a = m.forward_dat(rec, ranges, H, W)
        b = m.forward_events(*d, H, W)
        b2 = m(hb)
    assert torch.equal(b, b2), "the sampler is not repeatable on one histogram: %d" % int((b != b2).sum())
    if not torch.equal(a, b):        # diagnose: is one of them the FP32-pipe fall-back of algo="auto"?
        m.algo = "fp32"
        with torch.no_grad():
            f = m(hb)
        raise AssertionError("forward_dat vs forward_events: %d elements differ, max |d| %.3e; a == fp32 path: %s, "
                             "b == fp32 path: %s" % (int((a != b).sum()), float((a - b).abs().max()),
                                                     bool(torch.equal(a, f)), bool(torch.equal(b, f))))
