"""Oracle (TEST INFRASTRUCTURE): multi-step (parametric) LIF neuron, PyTorch restatement.

PARITY UNPINNED at this boundary: the arithmetic lives in the third-party package
``spikingjelly==0.0.0.0.14`` (``pip-requirements.txt:135``, ``conda-env.yml:369``,
``readme.md:17``), which is neither vendored under ``/root/reference`` nor installed.
This file restates the published algorithm of
``spikingjelly/activation_based/neuron.py`` (``BaseNode.single_step_forward`` /
``multi_step_forward``, ``ParametricLIFNode.neuronal_charge``, ``jit_soft_reset`` /
``jit_hard_reset``) and ``surrogate.py`` (``heaviside``, ``atan``, ``sigmoid``) and is
anchored on the reference's call site ``yolox/utils/utils_snn.py:44-53``:

    ParametricLIFNode(init_tau=2.0, decay_input=False, v_threshold=1.0, v_reset=None,
                      surrogate_function=ATan(alpha), detach_reset=False,
                      step_mode='m', backend='torch')

Per step (decay_input=False, v_reset in {None, 0}):
    v = v * (1 - sigmoid(w)) + x           # separate mul and add, as eager PyTorch does
    s = heaviside(v - v_th) = (v - v_th >= 0)
    v = v - s * v_th                       # soft reset (v_reset=None)
      | (1 - s) * v + s * v_reset          # hard reset
Surrogate gradients:  ATan    g * alpha / 2 / (1 + (pi/2 * alpha * x)^2)
                      Sigmoid g * alpha * sig(alpha x) * (1 - sig(alpha x))
                      Rect    g * 1[|x| < 0.5]      (yolox/models/activation.py:17-30; forward x > 0)
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn


def heaviside(x):
    return (x >= 0).to(x)


class _ATanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, alpha):
        ctx.save_for_backward(x)
        ctx.alpha = alpha
        return heaviside(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return ctx.alpha / 2 / (1 + (math.pi / 2 * ctx.alpha * x).pow(2)) * g, None


class _SigmoidFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, alpha):
        ctx.save_for_backward(x)
        ctx.alpha = alpha
        return heaviside(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        sg = (x * ctx.alpha).sigmoid()
        return g * (1.0 - sg) * sg * ctx.alpha, None


class ATan(nn.Module):
    def __init__(self, alpha=2.0, spiking=True):
        super().__init__()
        self.alpha, self.spiking = alpha, spiking

    def forward(self, x):
        return _ATanFn.apply(x, self.alpha)


class Sigmoid(nn.Module):
    def __init__(self, alpha=4.0, spiking=True):
        super().__init__()
        self.alpha, self.spiking = alpha, spiking

    def forward(self, x):
        return _SigmoidFn.apply(x, self.alpha)


def plif_forward(x_seq: torch.Tensor, w: torch.Tensor, surrogate: nn.Module,
                 v_threshold: float = 1.0, v_reset: Optional[float] = None,
                 decay_input: bool = False, detach_reset: bool = False,
                 v0: Optional[torch.Tensor] = None, return_v: bool = False):
    """Multi-step forward over ``x_seq[T, ...]``; returns the spike sequence (and v_seq)."""
    v = torch.zeros_like(x_seq[0]) if v0 is None else v0
    if v0 is None and v_reset is not None:
        v = v + v_reset
    out, vs = [], []
    sw = w.sigmoid()
    for t in range(x_seq.shape[0]):
        x = x_seq[t]
        if decay_input:
            if v_reset is None or v_reset == 0.0:
                v = v + (x - v) * sw
            else:
                v = v + (x - (v - v_reset)) * sw
        else:
            if v_reset is None or v_reset == 0.0:
                v = v * (1.0 - sw) + x
            else:
                v = v - (v - v_reset) * sw + x
        s = surrogate(v - v_threshold)
        sd = s.detach() if detach_reset else s
        if v_reset is None:
            v = v - sd * v_threshold
        else:
            v = (1.0 - sd) * v + sd * v_reset
        out.append(s)
        vs.append(v)
    if return_v:
        return torch.stack(out), torch.stack(vs)
    return torch.stack(out)


class OraclePLIF(nn.Module):
    """``neuron.ParametricLIFNode`` restated (attrs w, v, v_threshold, v_reset, ...)."""

    def __init__(self, init_tau=2.0, decay_input=True, v_threshold=1.0, v_reset=0.0,
                 surrogate_function=None, detach_reset=False, step_mode="s", backend="torch",
                 store_v_seq=False):
        super().__init__()
        self.w = nn.Parameter(torch.as_tensor(-math.log(init_tau - 1.0)))
        self.decay_input, self.v_threshold, self.v_reset = decay_input, v_threshold, v_reset
        self.surrogate_function = surrogate_function if surrogate_function is not None else Sigmoid()
        self.detach_reset, self.step_mode, self.backend = detach_reset, step_mode, backend
        self.store_v_seq = store_v_seq
        self.v = 0.0 if v_reset is None else v_reset

    def reset(self):
        self.v = 0.0 if self.v_reset is None else self.v_reset

    def forward(self, x):
        seq = x if self.step_mode == "m" else x.unsqueeze(0)
        v0 = self.v if isinstance(self.v, torch.Tensor) else torch.full_like(seq[0], float(self.v))
        s, vs = plif_forward(seq, self.w, self.surrogate_function, self.v_threshold, self.v_reset,
                             self.decay_input, self.detach_reset, v0=v0, return_v=True)
        self.v = vs[-1]
        if self.store_v_seq:
            self.v_seq = vs
        return s if self.step_mode == "m" else s[0]
