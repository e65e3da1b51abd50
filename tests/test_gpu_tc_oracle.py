"""The tcgen05 predictive engine (fp16 operands, fp32 accumulation) against the ORACLE at the shipped variational scales.

The headline bench runs `engine="tc"` at q_scale 1.351e-3 (experiment/ncmapss_lrt.yaml:22); ncmapss_fo ships 2.14e-4 with
S = 20 (experiment/ncmapss_fo.yaml:17-25) and ncmapss_rad 1.241e-3 with the radial guide.  These tests feed the oracle's
own weight noise (injected eps / r tensors; the oracle runs in float64) through bayesian.py:235-249 semantics on both sides
and bound every window's moments ELEMENT-WISE and RELATIVELY (no max-scaled absolute slack on ep_var).

Stated bound of the engine: pred / std / al_var within 1e-2 / 2e-2 / 2e-2 of the oracle at every scale (measured
< 2e-3); ep_var per window within a q_scale-DEPENDENT bound: a weight draw mu + sigma * eps is rounded to fp16 once (quantum
2^-11 |mu|, about 1.5e-5 for |mu| = 0.05), which acts as an extra, sample-independent weight noise of standard deviation
quantum / sqrt(3).  Against q_scale 1.351e-3 that is 0.6 % of the weight noise, against 2.14e-4 it is 4 % -- and with S
samples the CROSS term 2 * cov(loc, rounding) / var(loc) ~ 2 * (quantum / sigma) / sqrt(S) dominates the per-window error
(1.8 % median at 2.14e-4, S = 20), while the systematic inflation stays at (quantum / sigma)^2 (checked as `ep_bias`).
The radial guide spreads one sigma-sized step over a whole site (per-weight displacement sigma * |r| / sqrt(n), 3e-6 for the
153 600-element fc site): below the fp16 quantum, so radial ep_var on this engine is only good to tens of per cent -- stated,
tested at that bound, and irrelevant for `std` (ep_var is 1e-6 of al_var there); use engine="simt" when radial ep_var matters.
"""
import pytest
import torch

from oracle import bnn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NET = "inception"

# (q_scale, S, guide, ep_var bounds: per-window relative max, median over windows, |mean over windows of the ratio - 1|)
# measured on B200 (gpurun_out/tc_oracle_r2a.txt): lrt max 5.8e-3 / med 1.3e-3; fo max 8.3e-2 / med 1.8e-2; radial max 0.32 / med 0.098
CASES = [
    pytest.param(1.351e-3, 100, "normal", 2e-2, 5e-3, 5e-3, id="lrt_q1.351e-3_S100"),
    pytest.param(1.241e-3, 100, "radial", 6e-1, 2e-1, 2.5e-1, id="rad_q1.241e-3_S100"),
    pytest.param(2.14e-4, 20, "normal", 1.5e-1, 3e-2, 3e-2, id="fo_q2.14e-4_S20"),
]


@pytest.fixture(scope="module")
def eng():
    from bayesrul_b200 import Engine
    e = Engine(NET, DEV)
    assert e.has_tc()
    return e


def _rel(a, b):
    return ((a.double().cpu() - b.double()).abs() / b.double().abs().clamp_min(1e-300))


@pytest.mark.parametrize("q,S,guide,ep_max,ep_med,ep_bias", CASES)
def test_tc_moments_vs_oracle_shipped_scale(eng, q, S, guide, ep_max, ep_med, ep_bias):
    from bayesrul_b200 import Noise
    B = 512
    g = torch.Generator().manual_seed(int(q * 1e7) + S)
    x = torch.randn(B, 30, 18, generator=g)
    mu = O.init_params(NET, 12345)
    sg = torch.full_like(mu, q)
    P = mu.numel()
    eps = torch.randn(S, P, generator=g)
    nsites = len(O.site_table(NET))
    r = torch.randn(S, nsites, generator=g)
    noises = [O.InjectedNoise({"weight_eps": eps[s].double(), "radial_r": r[s].double()}) for s in range(S)]
    ref = O.predictive_moments(O.predict(NET, x.double(), mu.double(), sg.double(), guide, noises))
    nz = Noise(weight_eps=eps.to(DEV), radial_r=r.to(DEV) if guide == "radial" else None)
    got = eng.predict_moments(x.to(DEV), mu.to(DEV), sg.to(DEV), S=S, guide=guide, noise=nz, engine="tc")
    assert eng.tc_status() == 0
    names = ("pred", "std", "ep_var", "al_var")
    errs = {k: _rel(a, b) for k, a, b in zip(names, got, ref)}
    report = ", ".join(f"{k}: med {e.median().item():.2e} max {e.max().item():.2e}" for k, e in errs.items())
    print(f"\n[tc vs oracle] q={q} S={S} guide={guide}: {report}")
    assert errs["pred"].max().item() <= 1e-2, report
    assert errs["std"].max().item() <= 2e-2, report
    assert errs["al_var"].max().item() <= 2e-2, report
    assert errs["ep_var"].max().item() <= ep_max, report
    assert errs["ep_var"].median().item() <= ep_med, report
    bias = abs((got[2].double().cpu() / ref[2].double()).mean().item() - 1.0)  # systematic inflation of the epistemic variance
    print(f"[tc vs oracle] mean over windows of ep_var_tc / ep_var_oracle - 1 = {bias:.2e}")
    assert bias <= ep_bias, bias


def test_tc_mc_dropout_vs_oracle(eng):
    """configs[1] (ncmapss_mcd: p = 0.241437, 100 masks) on the fused engine against the oracle replaying the SAME Philox
    masks (PhiloxNoise.drop_mask): per-pass outputs and the predictive moments."""
    from bayesrul_b200 import Noise
    B, S, p = 256, 100, 0.241437
    g = torch.Generator().manual_seed(77)
    x = torch.randn(B, 30, 18, generator=g)
    mu = O.init_params(NET, 12345)
    ref_out = O.predict_mcd(NET, x, mu, p, [O.PhiloxNoise(NET, 4048, sample=s) for s in range(S)])
    ref = O.predictive_moments(ref_out.double())
    out = eng.forward(x.to(DEV), "det", theta=mu.to(DEV), S=S, p_dropout=p, noise=Noise(seed=4048), engine="tc")
    got = eng.predict_moments(x.to(DEV), mu.to(DEV), None, S=S, guide=None, p_dropout=p, noise=Noise(seed=4048), engine="tc")
    assert eng.tc_status() == 0
    e_out = _rel(out, ref_out)
    assert e_out.max().item() <= 1e-2, f"per-pass outputs: max rel err {e_out.max().item():.2e}"
    for k, a, b, tol in zip(("pred", "std", "ep_var", "al_var"), got, ref, (1e-2, 2e-2, 2e-2, 2e-2)):
        e = _rel(a, b)
        print(f"[tc mcd vs oracle] {k}: med {e.median().item():.2e} max {e.max().item():.2e}")
        assert e.max().item() <= tol, (k, e.max().item())
