"""Level-fused tcgen05 training kernels (brl_set_gemm_backend 'fused': csrc/brl_tc_train.cu) against the oracle.

One svi.step under `fit_ctxt` (bayesian.py:146-147; tyxe local_reparameterization / flipout restated in
oracle/bnn_oracle.py::forward_lrt / forward_flipout) with the oracle's own noise tensors injected, float64 autograd on the
oracle side.  Stated bound of this back-end (fp16 / bf16 operands, fp32 accumulation): outputs 1e-2, loss 5e-3, and for the
likelihood part of the gradient (the analytic KL gradient, identical on both sides, is subtracted so that it cannot mask
an error) per-site-group cosine > 0.999 and relative L2 error < 3e-2.  The fp32 FFMA back-end stays the 1e-3 parity engine
(tests/test_gpu_parity.py)."""
import pytest
import torch

from oracle import bnn_oracle as O
from tests.helpers import injected_to_engine, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NET = "inception"
KW = dict(guide="normal", prior_loc=0.0, dataset_size=238150)


@pytest.fixture(scope="module")
def eng():
    from bayesrul_b200 import Engine
    e = Engine(NET, DEV)
    yield e
    e.set_gemm_backend("simt")


def _kl_grads(mu, sg, prior_scale, c):
    return c * (mu / prior_scale**2), c * (-1.0 / sg + sg / prior_scale**2)


def _cos(a, b):
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-300))


def _check_grads(e, got, ref, mu, sg, prior_scale, tag, cos_min=0.999, rel_max=3e-2):
    c = 1.0 / (238150 * 540.0)
    kmu, ksg = _kl_grads(mu.double(), sg.double(), prior_scale, c)
    conv_end = e.info["sites"][20][0]  # sites 0..19 = the ten conv layers (fused kernels); 20..23 = fc + head (per-layer)
    for name, g, r, k in (("grad_mu", got["grad_mu"], ref["grad_mu"], kmu), ("grad_sigma", got["grad_sigma"], ref["grad_sigma"], ksg)):
        a = g.double().cpu() - k
        b = r.double() - k
        for part, sl in (("conv", slice(0, conv_end)), ("fc+head", slice(conv_end, None)), ("all", slice(None))):
            cs, rel = _cos(a[sl], b[sl]), float((a[sl] - b[sl]).norm() / (b[sl].norm() + 1e-300))
            print(f"[{tag}] {name}[{part}]: cosine {cs:.6f}, rel L2 err {rel:.2e}")
            assert cs > cos_min and rel < rel_max, (tag, name, part, cs, rel)
        # per conv layer (weights + bias of a layer form one group)
        for ly in range(10):
            lo, hi = e.info["sites"][2 * ly][0], e.info["sites"][2 * ly + 2][0]
            cs = _cos(a[lo:hi], b[lo:hi])
            assert cs > cos_min - 2e-3, (tag, name, "layer", ly, cs)


@pytest.mark.parametrize("mode,particles,sigma,prior_scale", [
    ("lrt", 1, 0.05, 0.138793), ("lrt", 1, 1.351e-3, 0.138793),
    ("flipout", 2, 0.05, 0.198768), ("flipout", 2, 2.14e-4, 0.198768)])
@pytest.mark.parametrize("B", [256, 37])
def test_fused_elbo_step_vs_oracle(eng, mode, particles, sigma, prior_scale, B):
    x, y, mu, sg = synth(NET, B, seed=21 + B, sigma=sigma)
    g = torch.Generator().manual_seed(4)
    nzs = [O.make_injected_noise(NET, B, mode, g) for _ in range(particles)]
    kw = dict(mode=mode, prior_scale=prior_scale, **KW)
    ref = O.elbo_loss_and_grads(NET, x.double(), y.double(), mu.double(), sg.double(),
                                noises=[O.InjectedNoise({k: v.double() for k, v in n.items()}) for n in nzs], **kw)
    eng.set_gemm_backend("fused")
    got = eng.elbo_step(x.to(DEV), y.to(DEV), mu.to(DEV), sg.to(DEV), particles=particles,
                        noise=injected_to_engine(NET, nzs, B, DEV), **kw)
    eng.set_gemm_backend("simt")
    assert eng.tc_status() == 0
    tag = f"{mode} sigma={sigma} B={B}"
    sc = got["scalars"].cpu()
    out_err = ((got["out"].cpu().double() - ref["out"]).abs() / ref["out"].abs().clamp_min(1e-3)).max().item()
    print(f"\n[{tag}] loss {sc[0].item():.6e} vs {ref['loss'].item():.6e}, nll {sc[1].item():.5e} vs {ref['nll_sum'].item():.5e}, "
          f"max rel out err {out_err:.2e}")
    assert abs(sc[0].item() / ref["loss"].item() - 1) < 5e-3
    assert abs(sc[1].item() / ref["nll_sum"].item() - 1) < 5e-3
    assert abs(sc[2].item() / ref["kl"].item() - 1) < 1e-4
    assert out_err < 1e-2
    # an entry-wise bound on the gradient of a ReLU net under operand rounding is not meaningful (a rounding flips the gate of
    # the pre-activations next to zero, and one flipped gate moves a unit's gradient by a whole window's contribution); the
    # direction is.  37 windows average fewer such flips than the 256 of a training minibatch, hence the wider small-batch bound
    # (measured worst case: LRT at sigma = 0.05, where eps * sqrt(var) is as large as the mean: cosine 0.9942)
    if B >= 256:
        _check_grads(eng, got, ref, mu, sg, prior_scale, tag)
    else:
        _check_grads(eng, got, ref, mu, sg, prior_scale, tag, cos_min=0.99, rel_max=1.5e-1)


@pytest.mark.parametrize("mode,particles", [("lrt", 1), ("flipout", 2)])
def test_fused_native_noise_matches_fp32_backend(eng, mode, particles):
    """Native Philox noise: the fused back-end draws the SAME eps / signs / weight noise as the fp32 back-end (same keys), so
    the two agree to the operand precision; and a graph-replayed step equals the eager one."""
    from bayesrul_b200 import Noise
    B = 256
    x, y, mu, sg = synth(NET, B, seed=5, sigma=0.02)
    x, y, mu, sg = x.to(DEV), y.to(DEV), mu.to(DEV), sg.to(DEV)
    kw = dict(mode=mode, prior_scale=0.138793, particles=particles, **KW)
    eng.set_gemm_backend("simt")
    ref = eng.elbo_step(x, y, mu, sg, noise=Noise(seed=11), **kw)
    eng.set_gemm_backend("fused")
    runs = [eng.elbo_step(x, y, mu, sg, noise=Noise(seed=11), **kw) for _ in range(3)]  # eager, captured, replayed
    other = eng.elbo_step(x, y, mu, sg, noise=Noise(seed=12), **kw)
    eng.set_gemm_backend("simt")
    assert eng.tc_status() == 0
    got = runs[0]
    assert abs(got["scalars"][0].item() / ref["scalars"][0].item() - 1) < 5e-3
    for k in ("grad_mu", "grad_sigma"):
        cs = _cos(got[k].double(), ref[k].double())
        assert cs > 0.999, (k, cs)
    for r in runs[1:]:  # atomics reorder the sums: equal up to fp32 summation order
        assert torch.allclose(r["scalars"], got["scalars"], rtol=1e-5)
        for k in ("grad_mu", "grad_sigma"):
            d = (r[k] - got[k]).abs().max().item() / got[k].abs().max().item()
            print(f"[{mode}] replay vs eager {k}: max |diff| / max |grad| = {d:.2e}")
            assert d < 1e-4, (k, d)
    assert abs(other["scalars"][1].item() - got["scalars"][1].item()) > 0  # a new seed is a new draw (the key is re-read)


@pytest.mark.parametrize("B", [1, 5, 129, 1000])
def test_fused_batch_sizes_match_fp32_backend(eng, B):
    """Ragged tiles (B not a multiple of 4), ragged fc M-tiles (not a multiple of 128) and several weight-gradient tiles per CTA
    (B = 1000: 250 tiles over <= 148 CTA groups): native noise, so the fp32 back-end draws the same numbers."""
    from bayesrul_b200 import Noise
    x, y, mu, sg = synth(NET, B, seed=100 + B, sigma=0.01)
    x, y, mu, sg = x.to(DEV), y.to(DEV), mu.to(DEV), sg.to(DEV)
    for mode, particles in (("lrt", 1), ("flipout", 2)):
        kw = dict(mode=mode, prior_scale=0.138793, particles=particles, **KW)
        eng.set_gemm_backend("simt")
        ref = eng.elbo_step(x, y, mu, sg, noise=Noise(seed=3), **kw)
        eng.set_gemm_backend("fused")
        got = eng.elbo_step(x, y, mu, sg, noise=Noise(seed=3), **kw)
        eng.set_gemm_backend("simt")
        assert eng.tc_status() == 0
        assert abs(got["scalars"][0].item() / ref["scalars"][0].item() - 1) < 5e-3, (mode, B)
        assert torch.allclose(got["out"], ref["out"], rtol=1e-2, atol=1e-3), (mode, B)
        c = 1.0 / (238150 * 540.0)
        kmu, ksg = _kl_grads(mu.double().cpu(), sg.double().cpu(), 0.138793, c)
        for k, kk in (("grad_mu", kmu), ("grad_sigma", ksg)):
            cs = _cos(got[k].double().cpu() - kk, ref[k].double().cpu() - kk)
            assert cs > (0.999 if B >= 129 else 0.98), (mode, B, k, cs)


@pytest.mark.parametrize("guide", ["normal", "radial"])
@pytest.mark.parametrize("B", [256, 33])
def test_fused_weight_sampling_elbo_vs_oracle(eng, guide, B):
    """Weight-sampling ELBO (no fit context / the radial guide's Trace_ELBO, bayesian.py:81-83,105-109) on the fused back-end:
    one contraction per layer with the particle's weight draw."""
    sigma = 0.03
    x, y, mu, sg = synth(NET, B, seed=70 + B, sigma=sigma)
    g = torch.Generator().manual_seed(8)
    nzs = [O.make_injected_noise(NET, B, "radial" if guide == "radial" else "ws", g)]
    kw = dict(mode="ws", guide=guide, prior_loc=0.0, prior_scale=0.138793, dataset_size=238150)
    ref = O.elbo_loss_and_grads(NET, x.double(), y.double(), mu.double(), sg.double(),
                                noises=[O.InjectedNoise({k: v.double() for k, v in n.items()}) for n in nzs], **kw)
    eng.set_gemm_backend("fused")
    got = eng.elbo_step(x.to(DEV), y.to(DEV), mu.to(DEV), sg.to(DEV), particles=1, noise=injected_to_engine(NET, nzs, B, DEV), **kw)
    eng.set_gemm_backend("simt")
    assert eng.tc_status() == 0
    sc = got["scalars"].cpu()
    assert abs(sc[0].item() / ref["loss"].item() - 1) < 5e-3
    assert abs(sc[1].item() / ref["nll_sum"].item() - 1) < 5e-3
    out_err = ((got["out"].cpu().double() - ref["out"]).abs() / ref["out"].abs().clamp_min(1e-3)).max().item()
    assert out_err < 1e-2, out_err
    cmin = 0.999 if B >= 256 else 0.99
    cs = _cos(got["grad_mu"].double().cpu(), ref["grad_mu"])  # the KL part of grad_mu is tiny next to the likelihood part here
    print(f"\n[ws {guide} B={B}] loss {sc[0].item():.6e} vs {ref['loss'].item():.6e}, out err {out_err:.2e}, grad_mu cosine {cs:.6f}")
    assert cs > cmin, cs


@pytest.mark.parametrize("p", [0.0, 0.241437])
@pytest.mark.parametrize("B", [256, 29])
def test_fused_hnn_step_vs_oracle(eng, p, B):
    """HNN.step (frequentist.py:39-48: forward with the dropout sites, F.gaussian_nll_loss, backward) on the fused back-end with the
    oracle's own keep masks injected."""
    x, y, mu, _ = synth(NET, B, seed=31 + B)
    g = torch.Generator().manual_seed(5)
    nz = O.make_injected_noise(NET, B, "det", g, p_dropout=p)
    th = mu.double().requires_grad_(True)
    loss, out = O.hnn_loss(NET, x.double(), y.double(), th, p, O.InjectedNoise({k: v.double() for k, v in nz.items()}))
    (gref,) = torch.autograd.grad(loss, th)
    eng.set_gemm_backend("fused")
    got = eng.hnn_step(x.to(DEV), y.to(DEV), mu.to(DEV), p, injected_to_engine(NET, [nz], B, DEV) if p > 0 else None)
    # native masks: same draws as the fp32 back-end, and graph replay == eager
    from bayesrul_b200 import Noise
    xd, yd, mud = x.to(DEV), y.to(DEV), mu.to(DEV)
    runs = [eng.hnn_step(xd, yd, mud, p, Noise(seed=9)) for _ in range(3)]
    eng.set_gemm_backend("simt")
    simt = eng.hnn_step(xd, yd, mud, p, Noise(seed=9))
    assert eng.tc_status() == 0
    assert abs(got["scalars"][0].item() / loss.item() - 1) < 5e-3
    out_err = ((got["out"].cpu().double() - out.detach()).abs() / out.detach().abs().clamp_min(1e-3)).max().item()
    cs = _cos(got["grad"].double().cpu(), gref)
    print(f"\n[hnn p={p} B={B}] loss {got['scalars'][0].item():.6e} vs {loss.item():.6e}, out err {out_err:.2e}, grad cosine {cs:.6f}")
    assert out_err < 1e-2
    assert cs > (0.999 if B >= 256 else 0.99)
    assert abs(runs[0]["scalars"][0].item() / simt["scalars"][0].item() - 1) < 5e-3
    assert _cos(runs[0]["grad"].double(), simt["grad"].double()) > 0.999
    for r in runs[1:]:
        assert torch.allclose(r["scalars"], runs[0]["scalars"], rtol=1e-5)
