"""torch.ops.bayesrul_b200.* (bayesrul_b200/ops.py): the registered operator set gives the Engine's numbers, has shape-only
(fake) implementations, and `elbo_loss(...).backward()` hands the fused step's gradients to autograd."""
import pytest
import torch

from oracle import bnn_oracle as O
from tests.helpers import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_registered_ops_match_engine_and_oracle():
    import bayesrul_b200.ops  # noqa: F401  (registers the ops)
    from bayesrul_b200 import Engine, Noise
    B, S = 48, 6
    x, y, mu, sg = synth("inception", B, seed=4, sigma=0.03)
    x, y, mu, sg = x.to(DEV), y.to(DEV), mu.to(DEV), sg.to(DEV)
    e = Engine("inception", DEV)
    want = e.predict_moments(x, mu, sg, S=S, guide="normal", noise=Noise(seed=7))
    got = torch.ops.bayesrul_b200.predict_moments(x, mu, sg, "inception", S, "normal", 0.0, 7, "simt")
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    empty = torch.empty(0, device=DEV)
    out = torch.ops.bayesrul_b200.forward(x, mu, empty, empty, "inception", "det", 1, 0.0, 0, "simt")
    ref = O.forward_det("inception", x.cpu(), mu.cpu())
    assert torch.allclose(out[0].cpu(), ref, rtol=1e-3, atol=1e-5)
    mm = torch.rand(5, 33, device=DEV) * 50, torch.rand(5, 33, device=DEV) + 0.5
    m, s = torch.ops.bayesrul_b200.mixture_moments(*mm)
    wm, ws = O.deep_ensemble_moments(mm[0].cpu(), mm[1].cpu())
    assert torch.allclose(m.cpu(), wm, rtol=1e-5) and torch.allclose(s.cpu(), ws, rtol=1e-4)
    # shape-only implementations (torch.library fake tensors)
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        fx = torch.empty(B, 30, 18, device=DEV)
        fm = torch.empty(mu.numel(), device=DEV)
        f = torch.ops.bayesrul_b200.predict_moments(fx, fm, fm, "inception", S, "normal", 0.0, 7, "simt")
        assert [tuple(t.shape) for t in f] == [(B,)] * 4


@pytest.mark.parametrize("backend", ["simt", "fused"])
def test_elbo_loss_is_differentiable(backend):
    from bayesrul_b200.ops import elbo_loss
    B = 64
    x, y, mu, sg = synth("inception", B, seed=9, sigma=0.02)
    x, y = x.to(DEV), y.to(DEV)
    mu = mu.to(DEV).requires_grad_(True)
    ls = sg.log().to(DEV).requires_grad_(True)
    kw = dict(net="inception", mode="lrt", guide="normal", particles=1, prior_loc=0.0, prior_scale=0.138793, dataset_size=238150, seed=3)
    loss = elbo_loss(mu, ls, x, y, backend=backend, **kw)
    (2.0 * loss).backward()
    ref = O.elbo_loss_and_grads("inception", x.cpu().double(), y.cpu().double(), mu.detach().cpu().double(), ls.detach().exp().cpu().double(),
                                mode="lrt", guide="normal", prior_loc=0.0, prior_scale=0.138793, dataset_size=238150,
                                noises=[O.PhiloxNoise("inception", 3)])
    tol = 1e-3 if backend == "simt" else 5e-3
    assert abs(loss.item() / ref["loss"].item() - 1) < tol
    for g, r in ((mu.grad, ref["grad_mu"]), (ls.grad, ref["grad_log_sigma"])):
        cs = float((g.cpu().double() * 2.0 * r).sum() / (g.cpu().double().norm() * (2.0 * r).norm()))
        assert cs > 0.999 and abs(float(g.cpu().double().norm() / (2.0 * r).norm()) - 1) < 2e-2, (backend, cs)
