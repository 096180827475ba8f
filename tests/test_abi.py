"""CPU: the C-ABI library loads, exports every symbol include/eas_b200.h declares, validates its
arguments on the host, and the Python product path refuses to run without CUDA (no fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

import eas_snn_b200 as eas
from eas_snn_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "eas_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(eas_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), n
        assert n in _lib.SIGNATURES, "python binding missing for %s" % n
    assert set(_lib.SIGNATURES) == set(names)
    assert L.eas_abi_version() == 1


def test_error_strings():
    L = _lib.lib()
    assert L.eas_error_string(0) == b"ok"
    for code in (-1, -2, -3, -4, -5):
        assert L.eas_error_string(code).startswith(b"EAS_E_")


def test_host_side_argument_checks_need_no_gpu():
    L = _lib.lib()
    assert L.eas_bin_events_ws_bytes(64, 4) >= 64 * 5 * 8
    # NULL pointers / bad shapes are rejected before anything is launched
    assert L.eas_bin_events(None, None, None, None, None, 2, 10, 8, 8, 4, None, None, 0, None) == -1
    assert L.eas_bin_events(None, None, None, None, None, 2, 10, 0, 8, 4, None, None, 0, None) == -2
    cfg = _lib.SamplerCfg(B=1, H=8, W=8, Tm=4, Ts=1, ksize=4, depth=2)
    assert L.eas_sampler_fwd(C.byref(cfg), None, None, None, None, None, None, 0, None) == -3
    cfg.ksize = 5
    assert L.eas_sampler_fwd_ws_bytes(C.byref(cfg)) > 0
    assert L.eas_sampler_fwd(C.byref(cfg), None, None, None, None, None, None, 0, None) == -1
    pc = _lib.PlifCfg(T=0, N=8)
    assert L.eas_plif_fwd(C.byref(pc), None, None, None, None, None, None) == -2


def test_host_side_checks_of_the_detector_and_representation_entry_points():
    """Shape / NULL / alignment contract violations come back as negative codes before anything is launched."""
    L = _lib.lib()
    E_NULL, E_SHAPE, E_UNSUP = -1, -2, -3
    assert L.eas_time_mean_planes(None, 3, 10, 12, 12, None, 12, 120, None) == E_SHAPE      # C % 8 != 0
    assert L.eas_time_mean_planes(None, 3, 10, 16, 16, None, 16, 160, None) == E_NULL
    assert L.eas_time_mean_planes(None, 3, 0, 16, 16, None, 16, 0, None) == 0               # nothing to do
    assert L.eas_upsample2x_planes(None, 2, 64, 1, 2, 2, 8, 4, None, 8, 256, None) == E_SHAPE   # in_ld < C
    assert L.eas_upsample2x_planes(None, 2, 64, 1, 2, 2, 8, 8, None, 8, 256, None) == E_NULL
    assert L.eas_yolox_decode(None, 1, 2, 3, 4, 4, 8.0, 1, None, 0, 6, None) == E_SHAPE       # n_ch < 5
    assert L.eas_yolox_decode(None, 1, 2, 3, 7, 7, 8.0, 1, None, 4, 6, None) == E_SHAPE       # anchors overflow
    assert L.eas_yolox_decode(None, 1, 2, 3, 7, 7, 8.0, 1, None, 0, 6, None) == E_NULL
    assert L.eas_focus_im2col(None, 1, 3, 4, None, 0, None) == E_SHAPE                        # odd height
    assert L.eas_focus_im2col(None, 1, 4, 4, None, 0, None) == E_NULL
    assert L.eas_hist_time_sum(None, 0, 2, 4, 10, None, None) == E_SHAPE                      # plane % 4 != 0
    assert L.eas_hist_time_sum(None, 2, 2, 4, 8, None, None) == E_UNSUP                       # bf16 histograms
    assert L.eas_hist_time_sum(None, 0, 2, 4, 8, None, None) == E_NULL
    assert L.eas_letterbox_bilinear(None, 0, 1, 4, 4, None, None, None, None, 8, 8, 0, 0, None, 8, 6, None) == E_SHAPE
    assert L.eas_letterbox_bilinear(None, 0, 1, 4, 4, None, None, None, None, 8, 8, 1, 0, None, 8, 8, None) == E_SHAPE
    assert L.eas_letterbox_bilinear(None, 0, 1, 4, 4, None, None, None, None, 8, 8, 0, 0, None, 8, 8, None) == E_NULL
    cc = _lib.ConvCfg(T=3, Tx=3, B=1, H=8, W=8, Cin=12, Cout=16, ksize=3, stride=1, n_wsplit=2, n_xsplit=1)
    assert L.eas_conv_bn_plif_fwd(C.byref(cc), None, None, None, None, None, None, 0, None) == E_SHAPE   # Cin % 8
    cc.Cin, cc.ksize = 16, 5
    assert L.eas_conv_bn_plif_fwd(C.byref(cc), None, None, None, None, None, None, 0, None) == E_UNSUP
    cc.ksize = 3
    assert L.eas_conv_bn_plif_fwd(C.byref(cc), None, None, None, None, None, None, 0, None) == E_NULL


def test_no_cpu_fallback():
    z16 = torch.zeros(4, dtype=torch.int16)
    with pytest.raises(_lib.EasError):
        eas.bin_events(z16, z16, torch.zeros(4, dtype=torch.int64), torch.zeros(4, dtype=torch.uint8),
                       torch.tensor([0, 4]), 4, 4, 2)
    m = eas.AdaptiveRSNNEmbedding(kernel_size=5, depth=2, nb_steps=4, thresh=1, vreset=0)
    with pytest.raises(_lib.EasError):
        m(torch.zeros(1, 4, 2, 8, 8))
    n = eas.ParametricLIFNode(step_mode="m", surrogate_function=eas.ATan(2.0), v_reset=None)
    with pytest.raises(_lib.EasError):
        n(torch.zeros(3, 2, 4))


def test_module_surface_matches_reference():
    """State-dict keys / ctor of the reference sampler (embedding.py:80-127) and PLIF (utils_snn.py:44-53)."""
    m = eas.AdaptiveRSNNEmbedding(kernel_size=5, in_channel=2, out_channel=2, readout="sum", split=False,
                                  write_zero=True, abs=False, depth=2, nb_steps=4, vreset=0, thresh=1,
                                  spike_fn=None, decay=None, embedding="arsnn", Ts=1, spike_attach=True)
    keys = set(m.state_dict())
    assert keys == {f"{s}_conv.{i}.{p}" for s in ("gate", "input") for i in (0, 2) for p in ("weight", "bias")}
    assert sum(p.numel() for p in m.parameters()) == 1216
    assert m(torch.zeros(2, 8, 8)).shape == (1, 2, 8, 8)             # 4-D/3-D passthrough broadcast (:144-146)
    n = eas.ParametricLIFNode(init_tau=2.0, decay_input=False, v_threshold=1.0, v_reset=None,
                              surrogate_function=eas.ATan(2.0), detach_reset=False, step_mode="m", backend="torch")
    assert list(n.state_dict()) == ["w"] and n.w.dim() == 0 and float(n.w) == 0.0
    assert n.v == 0.0 and eas.is_spiking_neuron(n)
    n.v = torch.ones(3)
    eas.reset_net(torch.nn.Sequential(n))
    assert n.v == 0.0


def test_read_dat_parses_header_and_records(tmp_path):
    """Host side of the .dat path: header lines, type/size bytes, raw records (dat_events_tools.py:126-181)."""
    import numpy as np
    from eas_snn_b200 import psee, synth
    from oracle import psee as opsee
    x, y, t, p = synth.dat_stream(seed=3, n=1000, H=24, W=32, span_us=10_000)
    path = tmp_path / "a_td.dat"
    with open(path, "wb") as f:
        f.write(b"% Data file containing Event2D events.\n% Version 2\n% Height 24\n% Width 32\n")
        f.write(bytes([0, 8]))
        opsee.pack_records(x, y, t, p).tofile(f)
    rec, (h, w) = psee.read_dat(str(path))
    assert (h, w) == (24, 32) and rec.shape == (1000, 2)
    assert np.array_equal(rec, psee.pack_records(x, y, t, p))
    dx, dy, dt, dp = opsee.decode(rec.view(opsee.EV_DTYPE).reshape(-1))
    assert np.array_equal(dx, x) and np.array_equal(dy, y) and np.array_equal(dt, t) and np.array_equal(dp, p)
