"""A context runs on ITS device whatever the caller's current device is, and creating / using it leaves the caller's
current device untouched (brl_create used to call cudaSetDevice and keep it)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _roundtrip(dev):
    from bayesrul_b200 import Engine, Noise
    from tests.helpers import synth
    x, _, mu, sg = synth("inception", 16, seed=3)
    e = Engine("inception", dev)
    x, mu, sg = x.to(dev), mu.to(dev), sg.to(dev)
    a = e.predict_moments(x, mu, sg, S=3, noise=Noise(seed=1))
    m = e.moments(e.forward(x, "det", theta=mu))
    y = torch.linspace(0, 90, 16, device=dev)
    e.set_gemm_backend("fused")  # the level-fused training kernels keep per-device state too
    r = e.elbo_step(x, y, mu, sg, mode="lrt", guide="normal", prior_loc=0.0, prior_scale=0.1, dataset_size=1000, noise=Noise(seed=2))
    e.set_gemm_backend("simt")
    assert e.tc_status() == 0
    return [t.cpu() for t in a], [t.cpu() for t in m] + [r["scalars"].float().cpu(), r["grad_mu"].cpu()]


def test_create_does_not_change_current_device():
    cur = torch.cuda.current_device()
    _roundtrip(f"cuda:{cur}")
    assert torch.cuda.current_device() == cur


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_engine_on_a_device_that_is_not_current():
    torch.cuda.set_device(0)
    ref = _roundtrip("cuda:0")
    got = _roundtrip("cuda:1")  # cuda:0 stays current throughout
    assert torch.cuda.current_device() == 0
    for u, v in zip(ref[0] + ref[1], got[0] + got[1]):
        assert torch.allclose(u, v, rtol=1e-4, atol=1e-6 * float(u.nan_to_num(0.0).abs().max()), equal_nan=True)  # gradients: atomics reorder sums
