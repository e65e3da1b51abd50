"""tcgen05 predictive engine of the Conv-D3 net (csrc/brl_tc_convd3.cuh; nets/conv.py:47-61, 73-77) vs the oracle and the fp32 engine.
fp16 operands, fp32 accumulation: the same stated bound as the Inception engine (outputs 1e-2 relative, moments 2e-2)."""
import pytest
import torch

from oracle import bnn_oracle as O
from tests.helpers import assert_close, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def eng():
    from bayesrul_b200 import Engine
    e = Engine("conv", DEV)
    assert e.has_tc()
    return e


@pytest.mark.parametrize("B", [1, 37, 128, 300])
def test_convd3_tc_det_forward_matches_oracle(eng, B):
    x, _, mu, _ = synth("conv", B, seed=B)
    ref = O.forward_det("conv", x, mu)
    out = eng.forward(x.to(DEV), "det", theta=mu.to(DEV), engine="tc")
    assert eng.tc_status() == 0
    assert_close(out[0], ref, rtol=1e-2, atol_scale=2e-3, what=f"conv tc det B={B}")


def test_convd3_tc_weight_samples_match_fp32_engine(eng):
    from bayesrul_b200 import Noise
    B, S = 200, 7
    x, _, mu, sg = synth("conv", B, seed=9, sigma=0.03)
    x, mu, sg = x.to(DEV), mu.to(DEV), sg.to(DEV)
    w = eng.sample_weights(mu, sg, "normal", S, Noise(seed=3))
    a = eng.forward(x, "ws", wsamp=w, S=S, engine="simt")
    b = eng.forward(x, "ws", wsamp=w, S=S, engine="tc")
    assert eng.tc_status() == 0
    assert_close(b, a, rtol=1e-2, atol_scale=2e-3, what="conv tc ws")


@pytest.mark.parametrize("guide", ["normal", "radial"])
def test_convd3_tc_predict_moments(eng, guide):
    """Same Philox weight draws on both engines (chunking differs on purpose: sample chunks merge by Chan's formula)."""
    from bayesrul_b200 import Noise
    B, S = 1000, 24
    x, _, mu, sg = synth("conv", B, seed=10, sigma=0.03)
    x, mu, sg = x.to(DEV), mu.to(DEV), sg.to(DEV)
    a = eng.predict_moments(x, mu, sg, S=S, guide=guide, noise=Noise(seed=8), engine="simt", chunk=6)
    b = eng.predict_moments(x, mu, sg, S=S, guide=guide, noise=Noise(seed=8), engine="tc", chunk=10)
    assert eng.tc_status() == 0
    for u, v, k in zip(b, a, ("pred", "std", "ep", "al")):
        if guide == "radial" and k == "ep":
            # the radial guide moves a weight by sigma * |r| / sqrt(n) of its site -- at or below the fp16 quantum of the weight -- so
            # the epistemic variance of this engine is only good to tens of per cent per window; std is unaffected
            assert_close(u, v, rtol=0.5, atol_scale=5e-2, what=f"{guide} {k}")
            assert abs(u.mean().item() / v.mean().item() - 1.0) < 0.1
        else:
            assert_close(u, v, rtol=2e-2, atol_scale=5e-3, what=f"{guide} {k}")


def test_convd3_tc_host_entry_point(eng):
    """brl_predict_moments_host on the Conv-D3 net: one copy + the device entry point on the tensor-core engine."""
    from bayesrul_b200 import Noise
    B, S = 700, 8
    x, _, mu, sg = synth("conv", B, seed=11, sigma=0.03)
    dev = eng.predict_moments(x.to(DEV), mu.to(DEV), sg.to(DEV), S=S, noise=Noise(seed=5), engine="tc")
    host = eng.predict_moments_host(x.pin_memory(), mu.to(DEV), sg.to(DEV), S=S, noise=Noise(seed=5), engine="tc")
    torch.cuda.synchronize()
    for i in range(4):
        assert torch.equal(host[i].to(DEV), dev[i]), i


def test_convd3_tc_rejects_dropout(eng):
    x, _, mu, _ = synth("conv", 8, seed=1)
    with pytest.raises(RuntimeError, match="dropout"):
        eng.forward(x.to(DEV), "det", theta=mu.to(DEV), S=2, p_dropout=0.2, engine="tc")


def test_convd3_tc_full_size_shard_invariance(eng):
    """B = 10 000 x S = 100 (the headline workload's size): the result of a window does not depend on how the batch is sharded (a weight
    draw has no window index, every tile row is computed independently) -- bit-exact, ragged split -- and the moments are finite."""
    from bayesrul_b200 import Noise
    B, S = 10000, 100
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, 30, 18, generator=g).to(DEV)
    mu = O.init_params("conv", 1).to(DEV)
    sg = torch.full_like(mu, 1.351e-3)
    full = eng.predict_moments(x, mu, sg, S=S, noise=Noise(seed=21), engine="tc")
    cut = 1003
    parts = [eng.predict_moments(x[a:b].contiguous(), mu, sg, S=S, noise=Noise(seed=21, window0=a), engine="tc") for a, b in ((0, cut), (cut, B))]
    assert eng.tc_status() == 0
    for i in range(4):
        assert torch.isfinite(full[i]).all()
        assert torch.equal(torch.cat([p[i] for p in parts]), full[i]), i
    assert (full[1] > 0).all() and (full[2] >= 0).all()
