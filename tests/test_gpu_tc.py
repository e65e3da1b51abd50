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


def test_tc_full_size_shard_properties(eng):
    """BASELINE size on the fused engine (B = 10 000 windows x S = 100): the result must not depend on how the windows are
    sharded over ranks (bit-exact: a weight draw has no window index, per-window noise is keyed by the global index) nor on
    how the MC samples are sharded (Chan merge of per-shard moments, dist.merge_moments), for weight sampling and MC-dropout."""
    from bayesrul_b200 import Noise
    from bayesrul_b200.dist import merge_moments, shard_range
    B, S = 10000, 100
    x, _, mu, sg = synth("inception", B, seed=78, sigma=0.02)
    x, mu, sg = x.to(DEV), mu.to(DEV), sg.to(DEV)
    for kw in (dict(sigma=sg, guide="normal"), dict(sigma=None, guide=None, p_dropout=0.241437)):
        full = eng.predict_moments(x, mu, S=S, noise=Noise(seed=21), engine="tc", **kw)
        # windows sharded over 3 "ranks" (ragged: 3334 / 3333 / 3333)
        parts = []
        for r in range(3):
            a, b = shard_range(B, r, 3)
            parts.append(eng.predict_moments(x[a:b].contiguous(), mu, S=S, noise=Noise(seed=21, window0=a), engine="tc", **kw))
        for i in range(4):
            assert torch.equal(torch.cat([p[i] for p in parts]), full[i]), (kw["guide"], i)
        # samples sharded over 4 "ranks" (25 each) + Chan merge
        sh = []
        for r in range(4):
            a, b = shard_range(S, r, 4)
            m = eng.predict_moments(x, mu, S=b - a, noise=Noise(seed=21, sample0=a), engine="tc", **kw)
            sh.append((b - a, m[0], m[2], m[3]))
        merged = merge_moments(sh)
        for u, v, k in zip(merged, full, ("pred", "std", "ep", "al")):
            assert_close(u, v, rtol=2e-4, atol_scale=1e-5, what=f"{kw['guide']} sample-sharded {k}")
        assert eng.tc_status() == 0


@pytest.mark.parametrize("B", [300, 5000])
@pytest.mark.parametrize("engine", ["tc", "simt"])
def test_host_entry_point_equals_device_entry_point(B, engine):
    """brl_predict_moments_host (host x, host results, window chunks overlapped with the copies on the fused engine) must
    give what brl_predict_moments gives on the device-resident batch: weight sampling (normal, radial) and MC-dropout."""
    from bayesrul_b200 import Engine, Noise
    e = Engine("inception", DEV)
    x, _, mu, sg = synth("inception", B, seed=79, sigma=0.02)
    xh = x.pin_memory()
    x, mu, sg = x.to(DEV), mu.to(DEV), sg.to(DEV)
    S = 7 if engine == "tc" else 3
    for kw in (dict(sigma=sg, guide="normal"), dict(sigma=sg, guide="radial"), dict(sigma=None, guide=None, p_dropout=0.241437)):
        ref = e.predict_moments(x, mu, S=S, noise=Noise(seed=5, window0=17), engine=engine, **kw)
        got = e.predict_moments_host(xh, mu, S=S, noise=Noise(seed=5, window0=17), engine=engine, **kw)
        assert got.shape == (4, B) and got.is_pinned()
        for i in range(4):
            assert_close(got[i], ref[i], rtol=1e-6, atol_scale=1e-7, what=f"{engine} {kw['guide']} output {i}")
    # pageable host memory works too (the copies are then synchronous)
    got2 = e.predict_moments_host(xh.clone(), mu, sg, S=S, guide="normal", noise=Noise(seed=5, window0=17), engine=engine)
    ref = e.predict_moments(x, mu, sg, S=S, guide="normal", noise=Noise(seed=5, window0=17), engine=engine)
    assert_close(got2[0], ref[0], rtol=1e-6, atol_scale=1e-7, what="pageable")
    with pytest.raises(RuntimeError):
        e.predict_moments_host(x, mu, sg, S=S, engine=engine)  # a device tensor is not a host batch


def test_tc_maximum_batch_and_limits(eng):
    """The largest batch one call accepts (B * 30 rows < 2^22: B <= 139 810) on the fused engine: first / last windows agree
    with the fp32 engine, one MC sample gives NaN epistemic variance like Tensor.var(0) of a single row (bayesian.py:148),
    and a larger batch is refused with an error, not truncated."""
    from bayesrul_b200 import Noise
    B = 139_810
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, 30, 18, generator=g).to(DEV)
    _, _, mu, sg = synth("inception", 4, seed=80, sigma=0.02)
    mu, sg = mu.to(DEV), sg.to(DEV)
    out = eng.forward(x, "det", theta=mu, engine="tc")[0]
    idx = torch.cat([torch.arange(0, 64), torch.arange(B - 64, B)]).to(DEV)
    ref = eng.forward(x[idx].contiguous(), "det", theta=mu, engine="simt")[0]
    assert_close(out[idx], ref, rtol=1e-2, atol_scale=2e-3, what="max batch det")
    pred, std, ep, al = eng.predict_moments(x, mu, sg, S=1, guide="normal", noise=Noise(seed=3), engine="tc")
    assert torch.isnan(ep).all() and torch.isfinite(pred).all() and (al > 0).all()
    assert eng.tc_status() == 0
    with pytest.raises(RuntimeError):
        eng.forward(torch.zeros(B + 1, 30, 18, device=DEV), "det", theta=mu, engine="tc")
