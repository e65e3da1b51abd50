"""tcgen05 TF32 per-layer GEMM back-end (brl_tc_gemm.cu) vs the fp32 SIMT back-end: same operators, same noise.
Stated bound (north_star: "bf16/TF32 tensor-core paths get a stated looser bound"): TF32 operands carry a 10-bit
mantissa with fp32 accumulation, so
  * every kernel class taken alone (forward outputs; input-gradient kernels and weight-gradient kernels fed by the SAME
    fp32 forward) agrees with the fp32 kernels to 5e-3 of the largest entry;
  * a full step on the tensor pipe agrees in loss to 5e-3, while its gradients are compared by direction / L2 norm:
    an operand rounding of 5e-4 flips the ReLU gate of the few pre-activations that sit within 5e-4 of zero, and one
    flipped gate moves that unit's gradient row by one window's whole contribution (~1/sqrt(rows) of the row, measured:
    tools/dbg_tf32.py), so an entry-wise bound on the gradient of a ReLU network is not meaningful."""
import pytest
import torch

from oracle import bnn_oracle as O
from tests.helpers import assert_close, injected_to_engine, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NETS = ["inception", "conv", "linear"]


@pytest.fixture(scope="module")
def engines():
    from bayesrul_b200 import Engine
    return {n: Engine(n, DEV) for n in NETS}


def _both(e, fn):
    e.set_gemm_backend("simt")
    a = fn()
    e.set_gemm_backend("tc")
    try:
        b = fn()
    finally:
        e.set_gemm_backend("simt")
    assert e.gemm_status() == 0
    return a, b


@pytest.mark.parametrize("net", NETS)
@pytest.mark.parametrize("B", [5, 130, 256])
def test_tf32_det_forward(engines, net, B):
    e = engines[net]
    x, _, mu, _ = synth(net, B, seed=B)
    a, b = _both(e, lambda: e.forward(x.to(DEV), "det", theta=mu.to(DEV)))
    assert_close(b, a, rtol=5e-3, atol_scale=1e-3, what=f"tf32 det {net} B={B}")
    assert_close(b[0], O.forward_det(net, x, mu), rtol=5e-3, atol_scale=1e-3, what=f"tf32 det vs oracle {net}")


@pytest.mark.parametrize("net", NETS)
@pytest.mark.parametrize("mode", ["lrt", "flipout"])
def test_tf32_forward_modes(engines, net, mode):
    from bayesrul_b200 import Noise
    e = engines[net]
    B = 70
    x, _, mu, sg = synth(net, B, seed=3, sigma=0.05)
    x, mu, sg = x.to(DEV), mu.to(DEV), sg.to(DEV)
    if mode == "lrt":
        fn = lambda: e.forward(x, "lrt", theta=mu, sigma=sg, noise=Noise(seed=11))
    else:
        w = e.sample_weights(mu, sg, "normal", 1, Noise(seed=5))
        fn = lambda: e.forward(x, "flipout", theta=mu, wsamp=w, noise=Noise(seed=11))
    a, b = _both(e, fn)
    assert_close(b, a, rtol=5e-3, atol_scale=2e-3, what=f"tf32 {mode} {net}")


STEP_CASES = [("lrt", "normal", 1), ("flipout", "normal", 2), ("ws", "normal", 1), ("ws", "radial", 1)]


def _step(e, net, mode, guide, particles, backend, B=96):
    from bayesrul_b200 import Noise
    x, y, mu, sg = synth(net, B, seed=21, sigma=0.05)
    x, y, mu, sg = x.to(DEV), y.to(DEV), mu.to(DEV), sg.to(DEV)
    kw = dict(mode=mode, guide=guide, particles=particles, prior_loc=0.0, prior_scale=0.138793, dataset_size=238150)
    e.set_gemm_backend(backend)
    try:
        r = e.elbo_step(x, y, mu, sg, noise=Noise(seed=77), **kw)
        torch.cuda.synchronize()
    finally:
        e.set_gemm_backend("simt")
    assert e.gemm_status() == 0
    return r


@pytest.mark.parametrize("net", NETS)
@pytest.mark.parametrize("mode,guide,particles", STEP_CASES)
@pytest.mark.parametrize("mask", [2, 4, 6])
def test_tf32_backward_kernels(engines, net, mode, guide, particles, mask):
    """input-gradient (2) / weight-gradient (4) kernels on the tensor pipe behind the fp32 forward: same ReLU gates,
    so the gradients must agree entry-wise."""
    e = engines[net]
    a = _step(e, net, mode, guide, particles, 0)
    b = _step(e, net, mode, guide, particles, mask)
    assert abs(b["scalars"][0].item() / a["scalars"][0].item() - 1) < 1e-6  # the forward is the same fp32 pass
    for k in ("grad_mu", "grad_sigma"):
        scale = a[k].abs().max().item()
        err = (a[k] - b[k]).abs().max().item()
        assert err <= 5e-3 * scale, (net, mode, mask, k, err, scale)


@pytest.mark.parametrize("net", NETS)
@pytest.mark.parametrize("mode,guide,particles", STEP_CASES)
def test_tf32_elbo_step(engines, net, mode, guide, particles):
    e = engines[net]
    a = _step(e, net, mode, guide, particles, "simt")
    b = _step(e, net, mode, guide, particles, "tc")
    assert abs(b["scalars"][0].item() / a["scalars"][0].item() - 1) < 5e-3
    assert_close(b["out"], a["out"], rtol=5e-3, atol_scale=2e-3, what=f"tf32 step out {net} {mode}")
    for k in ("grad_mu", "grad_sigma"):
        ga, gb = a[k].flatten().double(), b[k].flatten().double()
        cos = torch.nn.functional.cosine_similarity(ga, gb, dim=0).item()
        rel = ((ga - gb).norm() / ga.norm()).item()
        assert cos > 0.99 and rel < 0.15, (net, mode, k, cos, rel)


def test_tf32_lrt_step_vs_oracle(engines):
    """LRT ELBO step on the tensor pipe against autograd of the CPU oracle (injected noise)."""
    net, B = "inception", 64
    e = engines[net]
    g = torch.Generator().manual_seed(0)
    x, y, mu, sg = synth(net, B, seed=31, sigma=0.02)
    nz = O.make_injected_noise(net, B, "lrt", g)
    kw = dict(mode="lrt", guide="normal", prior_loc=0.0, prior_scale=0.138793, dataset_size=238150)
    ref = O.elbo_loss_and_grads(net, x.double(), y.double(), mu.double(), sg.double(),
                                noises=[O.InjectedNoise({k: v.double() for k, v in nz.items()})], **kw)
    e.set_gemm_backend("tc")
    try:
        got = e.elbo_step(x.to(DEV), y.to(DEV), mu.to(DEV), sg.to(DEV), noise=injected_to_engine(net, [nz], B, DEV), **kw)
    finally:
        e.set_gemm_backend("simt")
    assert e.gemm_status() == 0
    assert abs(got["scalars"][0].item() / ref["loss"].item() - 1) < 5e-3
    for k in ("grad_mu", "grad_sigma"):
        ga, gb = ref[k].flatten().double(), got[k].cpu().flatten().double()
        cos = torch.nn.functional.cosine_similarity(ga, gb, dim=0).item()
        assert cos > 0.995 and ((ga - gb).norm() / ga.norm()).item() < 0.1, (k, cos)
