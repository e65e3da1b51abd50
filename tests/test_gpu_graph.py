"""CUDA-graph replay of brl_elbo_step (native Philox noise) against the eager launches of the same step.

The replayed graph reads the per-step Philox key {seed, sample0, window0} from device memory and moves x / y / results
through staging buffers in the workspace, so it must give what the eager path gives for ANY sequence of seeds, window
offsets and input tensors (differences: fp32 atomic accumulation order only)."""
import pytest
import torch

from tests.helpers import assert_close, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

CASES = [("inception", "lrt", "normal", 1), ("inception", "flipout", "normal", 2), ("inception", "ws", "radial", 1),
         ("conv", "lrt", "normal", 2), ("linear", "flipout", "normal", 1), ("conv", "ws", "normal", 1)]


def _step(e, x, y, mu, sg, mode, guide, particles, seed, window0=0, grads=True):
    from bayesrul_b200 import Noise
    r = e.elbo_step(x, y, mu, sg, mode=mode, guide=guide, particles=particles, prior_loc=0.0, prior_scale=0.2,
                    dataset_size=238150, noise=Noise(seed=seed, window0=window0), compute_grads=grads)
    return {k: (v.clone() if v is not None else None) for k, v in r.items()}


@pytest.mark.parametrize("net,mode,guide,particles", CASES)
def test_graph_replay_matches_eager(net, mode, guide, particles):
    from bayesrul_b200 import Engine
    e = Engine(net, DEV)
    B = 96
    x0, y0, mu, sg = (t.to(DEV) for t in synth(net, B, seed=3, sigma=0.05))
    x1, y1, _, _ = (t.to(DEV) for t in synth(net, B, seed=4, sigma=0.05))
    plan = [(x0, y0, 11, 0), (x0, y0, 12, 0), (x1, y1, 13, 4096), (x0, y0, 11, 0), (x1, y1, 2 ** 40 + 5, 7)]
    e.set_step_graph(False)
    eager = [_step(e, x, y, mu, sg, mode, guide, particles, seed, w0) for x, y, seed, w0 in plan]
    e.set_step_graph(True)  # call 0 eager, call 1 captures + replays, calls 2.. replay
    replay = [_step(e, x, y, mu, sg, mode, guide, particles, seed, w0) for x, y, seed, w0 in plan]
    torch.cuda.synchronize()
    assert e.gemm_status() == 0
    for i, (a, b) in enumerate(zip(eager, replay)):
        assert_close(b["out"], a["out"], rtol=1e-5, what=f"step {i} out")
        assert_close(b["scalars"], a["scalars"], rtol=1e-5, what=f"step {i} scalars")
        # split-K forward kernels accumulate with fp32 atomics: a last-bit difference can flip the ReLU gate of a
        # pre-activation that sits at zero, which moves single gradient entries by one window's contribution
        for k in ("grad_mu", "grad_sigma", "grad_log_sigma"):
            assert_close(b[k], a[k], rtol=1e-3, atol_scale=1e-3, what=f"step {i} {k}")
    # the key really is per step: different seeds give different noise, the same seed the same
    assert not torch.allclose(replay[1]["out"], replay[0]["out"])
    assert_close(replay[3]["out"], replay[0]["out"], rtol=1e-6, what="same seed replayed")


def test_graph_replay_loss_only_and_injected_noise_stays_eager():
    from bayesrul_b200 import Engine, Noise
    e = Engine("inception", DEV)
    B = 64
    x, y, mu, sg = (t.to(DEV) for t in synth("inception", B, seed=5, sigma=0.05))
    e.set_step_graph(False)
    ref = _step(e, x, y, mu, sg, "lrt", "normal", 1, 77, grads=False)
    e.set_step_graph(True)
    for _ in range(3):
        got = _step(e, x, y, mu, sg, "lrt", "normal", 1, 77, grads=False)
    assert got["grad_mu"] is None
    assert_close(got["scalars"], ref["scalars"], rtol=1e-6, what="loss-only replay")
    # injected tensors: the eager path runs (a graph would freeze the injected pointers)
    eps = torch.randn(1, e.P, device=DEV)
    for _ in range(3):
        r = e.elbo_step(x, y, mu, sg, mode="ws", guide="normal", particles=1, prior_scale=0.2, dataset_size=238150,
                        noise=Noise(weight_eps=eps))
    w = mu + sg * eps[0]
    out = e.forward(x, "det", theta=w.contiguous())
    assert_close(r["out"][0], out[0], rtol=1e-5, what="injected weight draw")


@pytest.mark.parametrize("net,p", [("inception", 0.241437), ("conv", 0.0), ("linear", 0.3)])
def test_hnn_graph_replay_matches_eager(net, p):
    """brl_hnn_step (frequentist.py:39-48) replays a captured graph too: same losses, outputs and gradients as the eager launches
    for changing batches and (MC-dropout) mask keys."""
    from bayesrul_b200 import Engine, Noise
    e = Engine(net, DEV)
    B = 80
    x0, y0, th, _ = (t.to(DEV) for t in synth(net, B, seed=6))
    x1, y1, _, _ = (t.to(DEV) for t in synth(net, B, seed=7))
    plan = [(x0, y0, 1), (x0, y0, 2), (x1, y1, 3), (x0, y0, 1)]

    def run():
        return [{k: (v.clone() if v is not None else None) for k, v in
                 e.hnn_step(x, y, th, p_dropout=p, noise=Noise(seed=s, window0=5 * s)).items()} for x, y, s in plan]

    e.set_step_graph(False)
    eager = run()
    e.set_step_graph(True)
    replay = run()
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(zip(eager, replay)):
        assert_close(b["out"], a["out"], rtol=1e-5, what=f"step {i} out")
        assert_close(b["scalars"], a["scalars"], rtol=1e-5, what=f"step {i} scalars")
        assert_close(b["grad"], a["grad"], rtol=1e-3, atol_scale=1e-3, what=f"step {i} grad")
    if p > 0:
        assert not torch.allclose(replay[1]["out"], replay[0]["out"])  # another mask key, another mask
    assert_close(replay[3]["out"], replay[0]["out"], rtol=1e-6, what="same key replayed")
