"""ctypes binding of ``libeas_b200.so`` (C ABI in ``include/eas_b200.h``).

There is no CPU fallback: if the library is missing this module raises, and every op that
receives a non-CUDA tensor raises as well.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EAS_B200_LIB") or os.path.join(_HERE, "lib", "libeas_b200.so")

EAS_F32, EAS_I32, EAS_BF16, EAS_U8 = 0, 1, 2, 3
READOUT = {"sum": 0, "last": 1, "avg": 2}
SAMPLER_ALGO = {"auto": 0, "fp32": 1, "tensor": 2, "tensor_split": 3}
SURROGATE = {"atan": 0, "sigmoid": 1, "rect": 2}


class SamplerCfg(C.Structure):
    _fields_ = [("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Tm", C.c_int32), ("Ts", C.c_int32),
                ("ksize", C.c_int32), ("depth", C.c_int32), ("readout", C.c_int32), ("hard_reset", C.c_int32),
                ("vreset", C.c_float), ("thresh", C.c_float), ("spike_attach", C.c_int32),
                ("write_zero", C.c_int32), ("use_abs", C.c_int32), ("in_dtype", C.c_int32), ("algo", C.c_int32),
                ("surr_alpha", C.c_float)]


class SamplerPtrs(C.Structure):
    """``eas_sampler_weights`` / ``eas_sampler_grads`` (eight device pointers)."""
    _fields_ = [(n, C.c_void_p) for n in ("in_w0", "in_b0", "in_w1", "in_b1",
                                          "gate_w0", "gate_b0", "gate_w1", "gate_b1")]


class PlifCfg(C.Structure):
    _fields_ = [("T", C.c_int64), ("N", C.c_int64), ("v_threshold", C.c_float), ("hard_reset", C.c_int32),
                ("v_reset", C.c_float), ("decay_input", C.c_int32), ("detach_reset", C.c_int32),
                ("surrogate", C.c_int32), ("alpha", C.c_float), ("dtype", C.c_int32)]


class ConvCfg(C.Structure):
    _fields_ = [("T", C.c_int32), ("Tx", C.c_int32), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("Cin", C.c_int32), ("Cout", C.c_int32), ("ksize", C.c_int32), ("stride", C.c_int32),
                ("n_wsplit", C.c_int32), ("n_xsplit", C.c_int32), ("v_threshold", C.c_float),
                ("hard_reset", C.c_int32), ("v_reset", C.c_float), ("decay_input", C.c_int32),
                ("out_mode", C.c_int32), ("x_ld", C.c_int32), ("out_ld", C.c_int32), ("res_ld", C.c_int32),
                ("residual", C.c_void_p), ("w_unscale", C.c_void_p)]


# name -> (restype, argtypes); the single source the "exports every symbol" test walks.
_P = C.c_void_p
SIGNATURES = {
    "eas_abi_version": (C.c_int, []),
    "eas_error_string": (C.c_char_p, [C.c_int]),
    "eas_bin_events_ws_bytes": (C.c_size_t, [C.c_int64, C.c_int]),
    "eas_bin_events": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                 _P, _P, C.c_size_t, _P]),
    "eas_bin_events_ex": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                    _P, _P, C.c_size_t, _P, C.c_int, C.c_int]),
    "eas_dat_windows": (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.c_int64, C.c_int64, C.c_int32, _P, _P]),
    "eas_bin_dat_ws_bytes": (C.c_size_t, [C.c_int64, C.c_int]),
    "eas_bin_dat": (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.c_int, C.c_int, C.c_int, _P, _P, C.c_size_t, _P,
                              C.c_int, C.c_int]),
    "eas_hist_u8_bytes": (C.c_size_t, [C.c_int64, C.c_int, C.c_int, C.c_int]),
    "eas_hist_u8_expand": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P]),
    "eas_hist_u8_report": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, C.c_int, _P, _P]),
    "eas_hist_u8_set_sticky": (C.c_int, [_P]),
    "eas_sampler_fwd_ws_bytes": (C.c_size_t, [C.POINTER(SamplerCfg)]),
    "eas_sampler_fwd": (C.c_int, [C.POINTER(SamplerCfg), _P, C.POINTER(SamplerPtrs), _P, _P, _P, _P,
                                  C.c_size_t, _P]),
    "eas_sampler_bwd_ws_bytes": (C.c_size_t, [C.POINTER(SamplerCfg)]),
    "eas_sampler_bwd": (C.c_int, [C.POINTER(SamplerCfg), _P, C.POINTER(SamplerPtrs), _P, _P, _P,
                                  C.POINTER(SamplerPtrs), _P, _P, C.c_size_t, _P]),
    "eas_plif_fwd": (C.c_int, [C.POINTER(PlifCfg), _P, _P, _P, _P, _P, _P]),
    "eas_plif_bwd_ws_bytes": (C.c_size_t, [C.POINTER(PlifCfg)]),
    "eas_plif_bwd": (C.c_int, [C.POINTER(PlifCfg), _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "eas_conv_bn_plif_ws_bytes": (C.c_size_t, [C.POINTER(ConvCfg)]),
    "eas_conv_bn_plif_fwd": (C.c_int, [C.POINTER(ConvCfg), _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "eas_rvt_event_sum": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, C.c_int, _P, _P]),
    "eas_voxel_grid": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "eas_spp_pool_fwd": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "eas_time_mean_planes": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int, C.c_int, _P, C.c_int, C.c_int64, _P]),
    "eas_upsample2x_planes": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, _P,
                                        C.c_int, C.c_int64, _P]),
    "eas_letterbox_bilinear": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int, C.c_int, _P, _P, _P, _P, C.c_int, C.c_int,
                                         C.c_int, C.c_int, _P, C.c_int, C.c_int, _P]),
    "eas_hist_time_sum": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int, C.c_int64, _P, _P]),
    "eas_focus_im2col": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, _P, C.c_int64, _P]),
    "eas_yolox_decode": (C.c_int, [_P, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _P,
                                   C.c_int64, C.c_int64, _P]),
}

_lib = None


class EasError(RuntimeError):
    pass


def lib():
    """The loaded library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libeas_b200.so is not built (%s). Run `python -m eas_snn_b200.build` "
                "(or __graft_entry__.build()); there is no CPU fallback." % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().eas_error_string(rc)
        raise EasError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def ptr(t):
    """Device pointer of a tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise EasError("eas_snn_b200 ops need CUDA tensors: the sm_100a kernels are the only "
                           "implementation (no CPU fallback); got a %s tensor" % t.device)


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
