"""Multi-step parametric LIF neuron on sm_100a kernels.

Drop-in for ``spikingjelly.activation_based.neuron.ParametricLIFNode`` as the reference configures
it (``yolox/utils/utils_snn.py:44-53``): same constructor keywords, attributes (``w`` 0-dim
Parameter -> checkpoint key ``...act.w``, ``v``, ``v_threshold``, ``v_reset``,
``surrogate_function`` with ``.alpha``, ``detach_reset``, ``step_mode``, ``backend``) and
``reset()`` so that ``functional.reset_net`` finds it.  Forward and surrogate-gradient backward run
in ``eas_plif_fwd`` / ``eas_plif_bwd`` with the membrane potential in registers across all T steps.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib


class _Surrogate(nn.Module):
    """Configuration carrier (kind + alpha).  Calling it applies the Heaviside step with the
    surrogate gradient through plain torch ops -- plumbing for odd call sites, never the hot path
    (the neuron below reads ``kind``/``alpha`` and runs the fused kernels)."""

    kind = "atan"

    def __init__(self, alpha: float, spiking: bool = True):
        super().__init__()
        self.alpha = alpha
        self.spiking = spiking

    def extra_repr(self):
        return "alpha=%s" % self.alpha


class ATan(_Surrogate):
    kind = "atan"

    def __init__(self, alpha: float = 2.0, spiking: bool = True):
        super().__init__(alpha, spiking)


class Sigmoid(_Surrogate):
    kind = "sigmoid"

    def __init__(self, alpha: float = 4.0, spiking: bool = True):
        super().__init__(alpha, spiking)


class Rect(_Surrogate):
    """Rectangle window of yolox/models/activation.py:17-30 (alpha = 1)."""
    kind = "rect"

    def __init__(self, alpha: float = 1.0, spiking: bool = True):
        super().__init__(alpha, spiking)


def surrogate_kind(fn) -> tuple[int, float]:
    """Duck-typed: works for our carriers, for spikingjelly's ``surrogate.ATan/Sigmoid`` objects and for the
    reference's ``Rectangle`` autograd.Function CLASS (what ``spike_fn='rect'`` hands over,
    ``yolox/exp/event_yolox_base.py:146``; its ``alpha`` is a class attribute, activation.py:18)."""
    name = getattr(fn, "kind", None) or (fn.__name__ if isinstance(fn, type) else type(fn).__name__)
    name = name.lower()
    if name == "rectangle":
        name = "rect"
    if name not in _lib.SURROGATE:
        raise NotImplementedError("surrogate %r: kernels exist for ATan, Sigmoid, Rectangle" % name)
    alpha = getattr(fn, "alpha", 1.0)
    alpha = float(alpha.detach()) if isinstance(alpha, torch.Tensor) else float(alpha)
    return _lib.SURROGATE[name], alpha


def _plif_cfg(T, N, node, dtype):
    kind, alpha = surrogate_kind(node.surrogate_function)
    return _lib.PlifCfg(T=T, N=N, v_threshold=float(node.v_threshold),
                        hard_reset=0 if node.v_reset is None else 1,
                        v_reset=0.0 if node.v_reset is None else float(node.v_reset),
                        decay_input=int(bool(node.decay_input)), detach_reset=int(bool(node.detach_reset)),
                        surrogate=kind, alpha=alpha,
                        dtype=_lib.EAS_BF16 if dtype == torch.bfloat16 else _lib.EAS_F32)


class _PlifFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_seq, w, v0, node, want_v):
        L = _lib.lib()
        T = x_seq.shape[0]
        N = x_seq[0].numel()
        cfg = _plif_cfg(T, N, node, x_seq.dtype)
        spikes = torch.empty_like(x_seq)
        v_out = torch.empty(x_seq.shape[1:], dtype=torch.float32, device=x_seq.device) if want_v else None
        wd = w.detach().float().contiguous()
        with torch.cuda.device(x_seq.device):
            rc = L.eas_plif_fwd(C.byref(cfg), _lib.ptr(x_seq), _lib.ptr(wd), _lib.ptr(v0), _lib.ptr(spikes),
                                _lib.ptr(v_out), _lib.stream_ptr())
        _lib.check(rc, "eas_plif_fwd")
        ctx.cfg = cfg
        ctx.save_for_backward(x_seq, wd, v0)
        ctx.mark_non_differentiable(*([v_out] if v_out is not None else []))
        return spikes, v_out

    @staticmethod
    def backward(ctx, g_spikes, _g_v):
        L = _lib.lib()
        x_seq, wd, v0 = ctx.saved_tensors
        cfg = ctx.cfg
        g = g_spikes.contiguous().to(x_seq.dtype)
        gx = torch.empty_like(x_seq)
        gw = torch.empty((), dtype=torch.float32, device=x_seq.device)
        ws_bytes = L.eas_plif_bwd_ws_bytes(C.byref(cfg))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x_seq.device)
        with torch.cuda.device(x_seq.device):
            rc = L.eas_plif_bwd(C.byref(cfg), _lib.ptr(x_seq), _lib.ptr(wd), _lib.ptr(v0), _lib.ptr(g),
                                _lib.ptr(gx), _lib.ptr(gw), _lib.ptr(ws), ws_bytes, _lib.stream_ptr())
        _lib.check(rc, "eas_plif_bwd")
        return gx, gw, None, None, None


def plif_multistep(x_seq, w, node, v0=None, want_v=False):
    """Functional form: ``x_seq[T, ...]`` -> ``(spikes[T, ...], v_T or None)``."""
    _lib.require_cuda(x_seq, w)
    if x_seq.dtype not in (torch.float32, torch.bfloat16):
        x_seq = x_seq.float()
    if x_seq.dim() == 5 and not x_seq.is_contiguous() and x_seq.permute(0, 1, 3, 4, 2).is_contiguous():
        # [T, B, C, H, W] view of channels-last activations (what cuDNN's NHWC convolutions hand over): the neuron is
        # element-wise over everything but T, so it runs on the memory as it lies instead of forcing an NCHW copy
        v0p = None if v0 is None else v0.permute(0, 2, 3, 1)
        spikes, v_out = plif_multistep(x_seq.permute(0, 1, 3, 4, 2), w, node, v0=v0p, want_v=want_v)
        return spikes.permute(0, 1, 4, 2, 3), None if v_out is None else v_out.permute(0, 3, 1, 2)
    x_seq = x_seq.contiguous()
    if v0 is not None:
        # the kernels index v0[i] for every element of one time step: a state carried over from another batch
        # size / resolution (no reset_net in between) must not be read out of bounds or applied to other neurons
        if tuple(v0.shape) != tuple(x_seq.shape[1:]) or v0.device != x_seq.device:
            raise ValueError("membrane state of shape %s on %s does not match the input %s on %s: call reset() "
                             "(functional.reset_net) before changing the batch size or resolution"
                             % (tuple(v0.shape), v0.device, tuple(x_seq.shape[1:]), x_seq.device))
        v0 = v0.float().contiguous()
    return _PlifFn.apply(x_seq, w, v0, node, want_v)


class ParametricLIFNode(nn.Module):
    def __init__(self, init_tau: float = 2.0, decay_input: bool = True, v_threshold: float = 1.0,
                 v_reset: float | None = 0.0, surrogate_function: nn.Module | None = None,
                 detach_reset: bool = False, step_mode: str = "s", backend: str = "torch",
                 store_v_seq: bool = False):
        super().__init__()
        if store_v_seq:
            raise NotImplementedError("store_v_seq: the fused kernel never materialises v_seq")
        self.w = nn.Parameter(torch.as_tensor(-math.log(init_tau - 1.0)))
        self.decay_input = decay_input
        self.v_threshold = v_threshold
        self.v_reset = v_reset
        self.surrogate_function = surrogate_function if surrogate_function is not None else Sigmoid()
        self.detach_reset = detach_reset
        self.step_mode = step_mode
        self.backend = backend          # kept for signature parity; the sm_100a kernel is the backend
        self.store_v_seq = False
        self.keep_v = True              # write the final potential back (spikingjelly's stateful semantics)
        self.v = 0.0 if v_reset is None else v_reset

    def reset(self):
        self.v = 0.0 if self.v_reset is None else self.v_reset

    def extra_repr(self):
        return "v_threshold=%s, v_reset=%s, detach_reset=%s, step_mode=%s, backend=sm_100a" % (
            self.v_threshold, self.v_reset, self.detach_reset, self.step_mode)

    def forward(self, x):
        seq = x if self.step_mode == "m" else x.unsqueeze(0)
        v0 = self.v if isinstance(self.v, torch.Tensor) else None
        if v0 is not None and torch.is_grad_enabled() and x.requires_grad:
            # the carried potential is a plain buffer (grad_v0 is not produced): BPTT across separate calls inside a
            # network would be silently truncated.  The reference resets after every batch (trainer.py:115-117).
            # (A stand-alone node fed constants may carry its state with grad mode on: nothing flows back.)
            raise RuntimeError("ParametricLIFNode: state carried across calls under autograd is not differentiable here; "
                               "use step_mode='m' over the whole sequence and reset_net() between batches")
        spikes, v_out = plif_multistep(seq, self.w, self, v0=v0, want_v=self.keep_v)
        if self.keep_v:
            self.v = v_out
        return spikes if self.step_mode == "m" else spikes[0]


def is_spiking_neuron(module) -> bool:
    """``yolox/utils/utils_snn.py:12-13`` extended with the fused neuron."""
    return isinstance(module, ParametricLIFNode)


def reset_net(net: nn.Module):
    """``spikingjelly.activation_based.functional.reset_net``."""
    for m in net.modules():
        if hasattr(m, "reset"):
            m.reset()
