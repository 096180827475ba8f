"""GPU parity: eas_plif_fwd / eas_plif_bwd against the torch restatement of spikingjelly's
ParametricLIFNode (oracle/plif.py; PARITY UNPINNED at the spikingjelly boundary, see oracle/__init__)."""
import pytest
import torch

import eas_snn_b200 as eas
from oracle import plif as op
from helpers import close_report

pytestmark = pytest.mark.gpu

CASES = [
    # decay_input, v_reset, detach, surrogate, alpha, w
    (False, None, False, "atan", 2.0, 0.0),      # the reference configuration (utils_snn.py:44-53)
    (False, None, False, "atan", 2.0, 0.7),
    (False, None, True, "sigmoid", 4.0, -0.4),
    (True, 0.0, False, "atan", 1.5, 0.3),
    (True, 0.2, False, "sigmoid", 4.0, 0.3),
    (False, 0.2, False, "atan", 2.0, -0.2),
]


def _mk(kind, alpha, lib):
    if lib == "eas":
        return {"atan": eas.ATan, "sigmoid": eas.Sigmoid}[kind](alpha)
    return {"atan": op.ATan, "sigmoid": op.Sigmoid}[kind](alpha)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("shape", [(3, 2, 16, 12, 20), (4, 1, 3, 7, 5), (1, 1000)])
def test_fwd_bwd_fp32(cuda, case, shape):
    decay_input, v_reset, detach, kind, alpha, w0 = case
    g = torch.Generator().manual_seed(hash((case, shape)) % 2**31)
    x = torch.randn(shape, generator=g) * 1.2 + 0.3
    go = torch.randn(shape, generator=g)
    # oracle (CPU)
    xo = x.clone().requires_grad_(True)
    wo = torch.tensor(w0, requires_grad=True)
    so = op.plif_forward(xo, wo, _mk(kind, alpha, "op"), 1.0, v_reset, decay_input, detach)
    (so * go).sum().backward()
    # product (GPU)
    node = eas.ParametricLIFNode(init_tau=2.0, decay_input=decay_input, v_threshold=1.0, v_reset=v_reset,
                                 surrogate_function=_mk(kind, alpha, "eas"), detach_reset=detach,
                                 step_mode="m").to(cuda)
    node.w.data.fill_(w0)
    xg = x.to(cuda).requires_grad_(True)
    sg = node(xg)
    (sg * go.to(cuda)).sum().backward()
    mism = (sg.detach().cpu() != so.detach()).float().mean().item()
    assert mism <= 1e-4, "spike mismatch %.2e" % mism
    assert 0.02 < so.mean().item() < 0.98
    if mism == 0.0:
        ok, msg = close_report(xg.grad, xo.grad, rtol=1e-4, atol=1e-6)
        assert ok, "grad_x " + msg
        ok, msg = close_report(node.w.grad, wo.grad, rtol=1e-3, atol=1e-4)
        assert ok, "grad_w " + msg
    # final potential kept on the module, like spikingjelly
    _, vs = op.plif_forward(x, torch.tensor(w0), _mk(kind, alpha, "op"), 1.0, v_reset, decay_input, detach,
                            return_v=True)
    if mism == 0.0:
        ok, msg = close_report(node.v, vs[-1], rtol=1e-6, atol=1e-6)
        assert ok, "v_T " + msg


def test_default_w_is_bit_exact(cuda):
    """w = 0 -> decay 0.5 exactly; separate mul/add => potentials and spikes identical to torch."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn((3, 8, 64, 40, 48), generator=g)
    so, vo = op.plif_forward(x, torch.tensor(0.0), op.ATan(2.0), 1.0, None, False, False, return_v=True)
    node = eas.ParametricLIFNode(decay_input=False, v_reset=None, surrogate_function=eas.ATan(2.0),
                                 step_mode="m").to(cuda)
    sg = node(x.to(cuda))
    assert torch.equal(sg.cpu(), so)
    assert torch.equal(node.v.cpu(), vo[-1])


def test_state_carries_across_calls_until_reset(cuda):
    g = torch.Generator().manual_seed(4)
    x = torch.randn((6, 4, 33), generator=g) + 0.4
    node = eas.ParametricLIFNode(decay_input=False, v_reset=None, surrogate_function=eas.ATan(2.0),
                                 step_mode="m").to(cuda)
    full = node(x.to(cuda))
    node.reset()
    a = node(x[:3].to(cuda))
    b = node(x[3:].to(cuda))
    assert torch.equal(torch.cat([a, b]), full)
    node.reset()
    single = eas.ParametricLIFNode(decay_input=False, v_reset=None, surrogate_function=eas.ATan(2.0),
                                   step_mode="s").to(cuda)
    steps = torch.stack([single(x[t].to(cuda)) for t in range(6)])
    assert torch.equal(steps, full)


def test_bf16(cuda):
    g = torch.Generator().manual_seed(5)
    x = (torch.randn((3, 2, 8, 16, 16), generator=g) + 0.3).bfloat16()
    so = op.plif_forward(x.float(), torch.tensor(0.0), op.ATan(2.0), 1.0, None, False, False)
    node = eas.ParametricLIFNode(decay_input=False, v_reset=None, surrogate_function=eas.ATan(2.0),
                                 step_mode="m").to(cuda)
    sg = node(x.to(cuda))
    assert sg.dtype == torch.bfloat16
    assert torch.equal(sg.float().cpu(), so)      # bf16 inputs are exact in fp32; state is fp32 in registers


def test_large_backbone_shape_property(cuda):
    """SYOLOX-M dark2-sized tensor: T=3, 8 x 96 x 64 x 80 (no oracle needed: recurrence identities)."""
    T, shape = 3, (8, 96, 64, 80)
    x = torch.rand((T,) + shape, device=cuda) * 1.5
    node = eas.ParametricLIFNode(decay_input=False, v_reset=None, surrogate_function=eas.ATan(2.0),
                                 step_mode="m").to(cuda)
    s = node(x)
    # soft reset identity: sum_t (0.5^(T-1-t)) x_t - sum_t (0.5^(T-1-t)) s_t == v_T  (w=0 -> decay 0.5)
    wts = torch.tensor([0.5 ** (T - 1 - t) for t in range(T)], device=cuda).view(T, 1, 1, 1, 1)
    v_T = ((x - s) * wts).sum(0)
    assert torch.allclose(node.v, v_T, atol=1e-5)
    assert set(s.unique().tolist()) <= {0.0, 1.0}


def test_channels_last_views_run_in_place_and_match_the_contiguous_path(cuda):
    """[T, B, C, H, W] views of channels-last activations (cuDNN's NHWC convolutions in the training path) are
    processed on the memory as it lies: same spikes, state and input gradients as the NCHW-contiguous copy, bit for bit, and
    the outputs keep the channels-last layout (no conversion kernels between conv and neuron)."""
    torch.manual_seed(9)
    T, B, Cc, H, W = 3, 2, 8, 6, 10
    base = (torch.randn(T * B, Cc, H, W, device=cuda) + 0.4).contiguous(memory_format=torch.channels_last)
    outs = []
    for cl in (True, False):
        x = (base if cl else base.contiguous()).detach().clone(memory_format=torch.preserve_format).requires_grad_(True)
        node = eas.ParametricLIFNode(decay_input=False, v_reset=None, surrogate_function=eas.ATan(2.0),
                                     step_mode="m").to(cuda)
        s = node(x.view(T, B, Cc, H, W))
        if cl:
            assert s.permute(0, 1, 3, 4, 2).is_contiguous() and node.v.permute(0, 2, 3, 1).is_contiguous()
        g = torch.arange(s.numel(), device=cuda, dtype=torch.float32).view_as(s).sin()
        (s * g).sum().backward()
        outs.append((s.detach().clone(), node.v.clone(), x.grad.clone(), node.w.grad.clone()))
    for a, b in list(zip(*outs))[:3]:
        assert torch.equal(a.contiguous(), b.contiguous())
    gw_cl, gw = outs[0][3], outs[1][3]                  # d w: same terms, summed in memory order
    assert abs(float(gw_cl) - float(gw)) <= 1e-5 * abs(float(gw)) + 1e-7
