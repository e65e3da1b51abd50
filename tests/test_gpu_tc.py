"""tcgen05 engine (fp16 operands, fp32 accumulation) vs the fp32 SIMT engine and the oracle.
Stated bound for this engine: outputs within 1e-2 relative (fp16 has a 10-bit mantissa, the same as
TF32; every layer's operands are rounded once), predictive moments within 2e-2."""
import pytest
import torch

from oracle import bnn_oracle as O
from tests.helpers import assert_close, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def eng():
    from bayesrul_b200 import Engine
    e = Engine("inception", DEV)
    assert e.has_tc()
    return e


@pytest.mark.parametrize("B", [4, 37, 128, 300])
def test_tc_det_forward(eng, B):
    x, _, mu, _ = synth("inception", B, seed=B)
    ref = O.forward_det("inception", x, mu)
    out = eng.forward(x.to(DEV), "det", theta=mu.to(DEV), engine="tc")
    assert eng.tc_status() == 0
    assert_close(out[0], ref, rtol=1e-2, atol_scale=2e-3, what=f"tc det B={B}")


def test_tc_weight_samples_match_simt(eng):
    from bayesrul_b200 import Noise
    B, S = 70, 5
    x, _, mu, sg = synth("inception", B, seed=9, sigma=0.03)
    x, mu, sg = x.to(DEV), mu.to(DEV), sg.to(DEV)
    w = eng.sample_weights(mu, sg, "normal", S, Noise(seed=3))
    a = eng.forward(x, "ws", wsamp=w, S=S, engine="simt")
    b = eng.forward(x, "ws", wsamp=w, S=S, engine="tc")
    assert eng.tc_status() == 0
    assert_close(b, a, rtol=1e-2, atol_scale=2e-3, what="tc ws")


def test_tc_predict_moments(eng):
    from bayesrul_b200 import Noise
    B, S = 500, 12
    x, _, mu, sg = synth("inception", B, seed=10, sigma=0.03)
    x, mu, sg = x.to(DEV), mu.to(DEV), sg.to(DEV)
    a = eng.predict_moments(x, mu, sg, S=S, noise=Noise(seed=8), engine="simt", chunk=4)
    b = eng.predict_moments(x, mu, sg, S=S, noise=Noise(seed=8), engine="tc", chunk=5)
    assert eng.tc_status() == 0
    for u, v, k in zip(b, a, ("pred", "std", "ep", "al")):
        assert_close(u, v, rtol=2e-2, atol_scale=5e-3, what=k)


def test_tc_mc_dropout(eng):
    from bayesrul_b200 import Noise
    B, S, p = 64, 3, 0.241437
    x, _, mu, _ = synth("inception", B, seed=12)
    x, mu = x.to(DEV), mu.to(DEV)
    a = eng.forward(x, "det", theta=mu, S=S, p_dropout=p, noise=Noise(seed=77), engine="simt")
    b = eng.forward(x, "det", theta=mu, S=S, p_dropout=p, noise=Noise(seed=77), engine="tc")
    assert eng.tc_status() == 0
    assert_close(b, a, rtol=1e-2, atol_scale=2e-3, what="tc mcd")
