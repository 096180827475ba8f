"""GPU parity: sampler backward (BPTT, Rectangle surrogate, SAT / RPD flags) against the golden
parameter gradients produced by the reference's autograd, and against autograd of the dense oracle
(including the gradient w.r.t. the input micro-bins)."""
import pytest
import torch

import eas_snn_b200 as eas
from oracle import sampler as osamp
from helpers import load_golden, sampler_case, sampler_kwargs

pytestmark = pytest.mark.gpu

NAMES = list(load_golden("sampler")["names"])


def _grad_close(got, want, name):
    """Gradients are O(1e2-1e4) sums over ~1e4 pixels: compare relative to the tensor's scale."""
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    scale = want.abs().max().clamp(min=1e-6)
    err = (got - want).abs().max() / scale
    return float(err), "%s: max err / max|g| = %.3e (max|g| %.3e)" % (name, float(err), float(scale))


@pytest.mark.parametrize("name", NAMES)
def test_param_grads_golden(cuda, name):
    z = load_golden("sampler")
    cfg, params, grads, x, y = sampler_case(z, name)
    m = eas.AdaptiveRSNNEmbedding(**sampler_kwargs(cfg)).to(cuda)
    m.load_state_dict(params)
    out = m(x.to(cuda))
    wgt = torch.linspace(-1.0, 1.0, out.numel(), device=cuda).view_as(out)
    got = torch.autograd.grad((out * wgt).sum(), list(m.parameters()))
    worst, msgs = 0.0, []
    for (pn, _), g in zip(m.named_parameters(), got):
        err, msg = _grad_close(g, grads[pn], pn)
        worst = max(worst, err)
        msgs.append(msg)
    print(name, "\n  " + "\n  ".join(msgs))
    assert worst < 2e-3, name + ": " + "; ".join(msgs)


def test_input_and_param_grads_vs_oracle_autograd(cuda):
    torch.manual_seed(80)
    kw = dict(kernel_size=5, in_channel=2, out_channel=2, readout="sum", split=False, write_zero=True, abs=False,
              depth=2, nb_steps=4, vreset=0, thresh=1, embedding="arsnn", Ts=1, spike_attach=True)
    ref = osamp.OracleSampler(**kw)
    g = torch.Generator().manual_seed(7)
    x = torch.poisson(torch.full((2, 4, 2, 48, 80), 1.3), generator=g)
    go = torch.randn((1, 2, 2, 48, 80), generator=g)
    xr = x.clone().requires_grad_(True)
    (ref(xr) * go).sum().backward()
    m = eas.AdaptiveRSNNEmbedding(**kw).to(cuda)
    m.load_state_dict(ref.state_dict())
    xg = x.to(cuda).requires_grad_(True)
    (m(xg) * go.to(cuda)).sum().backward()
    err, msg = _grad_close(xg.grad, xr.grad, "d events")
    print(msg)
    assert err < 2e-3, msg
    for (pn, pr), (_, pg) in zip(ref.named_parameters(), m.named_parameters()):
        err, msg = _grad_close(pg.grad, pr.grad, pn)
        print(msg)
        assert err < 2e-3, msg


def test_training_step_changes_parameters(cuda):
    """One optimiser step through the module API (what trainer.py:104-114 does)."""
    torch.manual_seed(0)
    m = eas.AdaptiveRSNNEmbedding(kernel_size=5, depth=2, nb_steps=4, thresh=1, vreset=0, Ts=1, write_zero=True,
                                  spike_attach=True).to(cuda)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    x = torch.poisson(torch.full((2, 4, 2, 32, 64), 1.5)).to(cuda)
    before = [p.detach().clone() for p in m.parameters()]
    loss = m(x).square().mean()
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    opt.step()
    assert any(not torch.equal(a, b) for a, b in zip(before, m.parameters()))
