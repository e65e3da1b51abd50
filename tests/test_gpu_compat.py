"""Reference-facing mirrors (BNN / HNN / shims) end to end on the GPU: same hooks, keys and returns as
bayesrul/models/bayesian.py and frequentist.py."""
import numpy as np
import pytest
import torch

from oracle import bnn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _batch(B, seed=0, device=DEV):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 30, 18, generator=g).to(device), (torch.rand(B, generator=g) * 100).to(device)


@pytest.mark.parametrize("fit_context,guide,particles", [("lrt", "normal", 1), ("flipout", "normal", 2), (None, "radial", 1),
                                                          (None, "normal", 1)])
def test_bnn_training_loop_and_checkpoint(fit_context, guide, particles):
    from bayesrul_b200.compat import BNN, Inception, pyro_shim
    torch.manual_seed(12345)
    net = Inception(30, 18)
    with torch.no_grad():
        net.last.bias.fill_(3.0)  # keep the aggregated scale above ln 2 so that inverse_softplus stays positive (A.5)
    opt = pyro_shim.ClippedAdam({"lr": 0.000857, "betas": [0.95, 0.999], "clip_norm": 15})
    m = BNN(net, opt, pretrain_epochs=5, mc_samples_train=particles, mc_samples_eval=8, dataset_size=238150,
            fit_context=fit_context, prior_loc=0.0, prior_scale=0.138793, guide=guide, q_scale=1.351e-3, device=DEV)
    m.on_fit_start()
    x, y = _batch(100)
    loc0 = m.bnn.net_guide.loc.clone()
    elbos = []
    for i in range(5):
        m.training_step((x, y), i)
        elbos.append(m.logged["elbo/train"])
    assert set(m.logged) >= {"mse/train", "elbo/train", "kl/train", "likelihood/train", "rmsce/train", "sharp/train"}
    assert all(np.isfinite(e) for e in elbos) and elbos[-1] < elbos[0]
    # SURVEY A.10 anchor: KL/(N*540) ~ 0.006 for the LRT hyper-parameters; elbo of that order
    assert 0.004 < m.logged["kl/train"] / (238150 * 540) < 0.008
    assert not torch.equal(loc0, m.bnn.net_guide.loc)
    m.validation_step((x, y), 0)
    assert np.isfinite(m.logged["elbo/val"]) and set(m.logged) >= {"mse/val", "kl/val", "rmsce/val", "sharp/val"}
    nll = m.test_step((x, y), 0)
    assert np.isfinite(float(nll))
    p = m.predict_step((x.cpu(), y.cpu()), 0)
    assert set(p) == {"labels", "ep_vars", "al_vars", "preds", "stds"}
    assert all(isinstance(v, np.ndarray) and v.shape == (100,) for v in p.values())
    np.testing.assert_allclose(p["stds"] ** 2, p["ep_vars"] + p["al_vars"], rtol=1e-4)
    # checkpoint round trip through the pyro-style param store
    ck = {}
    m.on_save_checkpoint(ck)
    names = list(ck["param_store"]["params"])
    assert len(names) == 48 and names[0] == "net_guide.net.layers.0.conv1.0.weight.loc"
    saved = {k: v.clone() for k, v in ck["param_store"]["params"].items()}
    m2 = BNN(Inception(30, 18), None, 5, particles, 8, 238150, fit_context, 0.0, 0.138793, guide, 1.351e-3, device=DEV)
    m2.on_load_checkpoint({"param_store": {"params": saved, "constraints": ck["param_store"]["constraints"]}, "state_dict": {}})
    m2.on_predict_start()
    assert torch.equal(m2.bnn.net_guide.loc, m.bnn.net_guide.loc)
    assert torch.allclose(m2.bnn.net_guide.scale, m.bnn.net_guide.scale)


@pytest.mark.parametrize("train_backend,tol", [("simt", 2e-3), ("auto", 5e-3)])
def test_bnn_step_matches_oracle_semantics(train_backend, tol):
    """svi.step's return value is the scaled loss of SURVEY A.6 (checked with the oracle on the same Philox draw), on the fp32
    parity back-end and on the default one ("auto": the level-fused tcgen05 kernels for Inception under LRT)."""
    from bayesrul_b200.compat import BNN, Inception
    torch.manual_seed(7)
    m = BNN(Inception(30, 18), None, 5, 1, 8, 238150, "lrt", 0.0, 0.138793, "normal", 0.02, device=DEV, train_backend=train_backend)
    m.on_fit_start()
    x, y = _batch(64, 3)
    g = m.bnn.net_guide
    mu, sg = g.loc.cpu().clone(), g.scale.cpu().clone()
    step = m.bnn._step + 1
    seed = (m.bnn.seed + 0x9E3779B97F4A7C15 * step) & 0xFFFFFFFFFFFFFFFF
    with m.fit_ctxt():
        loss = m.svi.step(x, y.unsqueeze(-1))
    ref, _ = O.elbo_loss("inception", x.cpu(), y.cpu(), mu, sg, mode="lrt", guide="normal", prior_loc=0.0,
                         prior_scale=0.138793, dataset_size=238150, noises=[O.PhiloxNoise("inception", seed)])
    assert abs(loss / ref.item() - 1) < tol


def test_hnn_train_and_mc_dropout_predict():
    from bayesrul_b200.compat import HNN, Inception
    from functools import partial
    torch.manual_seed(1)
    net = Inception(30, 18, dropout=0.241437)
    m = HNN(net, partial(torch.optim.Adam, lr=1e-3), mc_samples=10, p_dropout=0.241437, device=DEV)
    opt = m.configure_optimizers()
    x, y = _batch(100, 5)
    losses = []
    for i in range(8):  # Lightning's automatic-optimisation closure order: training_step, zero_grad, backward, step
        m.net.train()
        loss = m.training_step((x, y), i)
        assert loss.requires_grad and loss.grad_fn is not None
        opt.zero_grad()
        loss.backward()
        assert all(p.grad is not None and p.grad.shape == p.shape for p in m.net.parameters())
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0]
    m.net.eval()
    p = m.predict_step((x.cpu(), y.cpu()), 0)
    assert set(p) == {"labels", "ep_vars", "al_vars", "preds", "stds"} and p["preds"].shape == (100,)
    assert (p["ep_vars"] > 0).all()
    # a host batch goes through brl_predict_moments_host, a device batch through brl_predict_moments: same masks, same numbers
    it0 = m._it
    p_host = m.predict_step((x.cpu(), y.cpu()), 0)
    m._it = it0
    p_dev = m.predict_step((x, y), 0)
    for k in ("preds", "stds", "ep_vars", "al_vars"):
        np.testing.assert_allclose(p_host[k], p_dev[k], rtol=1e-5, atol=1e-6)
    out = m.validation_step((x, y), 0)
    assert set(out) == {"loss", "label", "pred", "std"}
    m.test_step((x, y), 0)
    assert np.isfinite(float(m.logged["nll/test"]))
    # deterministic HNN (dropout 0): eval forward == oracle
    net0 = Inception(30, 18).to(DEV)
    ref = O.forward_det("inception", x.cpu(), net0.flat().cpu())
    assert torch.allclose(net0(x).cpu(), ref, rtol=1e-3, atol=1e-5)


def test_deep_ensemble_dataframe(golden_dir):
    import os
    import pandas as pd
    from bayesrul_b200.compat import deep_ensemble
    z = np.load(os.path.join(golden_dir, "deep_ensemble.npz"))
    M, n = z["mu_m"].shape
    df = pd.concat([pd.DataFrame(dict(model=f"HNN_{k:03d}", preds=z["mu_m"][k], stds=z["sigma_m"][k], labels=np.zeros(n)))
                    for k in range(M)])
    de = deep_ensemble(df)
    np.testing.assert_allclose(de.preds.values, z["preds"], rtol=1e-5)
    np.testing.assert_allclose(de.stds.values, z["stds"], rtol=1e-3, atol=1e-4)


def test_shims_installable():
    from bayesrul_b200.compat import install_shims
    install_shims()
    import pyro
    import tyxe
    from pyro.infer import SVI, Trace_ELBO, TraceMeanField_ELBO  # noqa: F401
    assert callable(tyxe.poutine.local_reparameterization) and callable(tyxe.poutine.flipout)
    pyro.clear_param_store()
    assert pyro.get_param_store().get_state()["params"] == {}


@pytest.mark.parametrize("engine", ["simt", "tc"])
def test_chunked_host_predict_equals_device_predict(engine):
    """predict_moments_host overlaps chunked copies of a pinned HOST batch with the compute; the chunks see the same
    weight draws (the Philox key of a draw has no window index), so the result equals the one-shot device call."""
    from bayesrul_b200.compat import BNN, Inception
    m = BNN(Inception(30, 18), None, 0, 1, 6, 238150, "lrt", 0.0, 0.138793, "normal", 0.05, device=DEV, engine=engine)
    m.on_predict_start()
    g = torch.Generator().manual_seed(9)
    x = torch.randn(3000, 30, 18, generator=g).pin_memory()
    y = torch.rand(3000, generator=g)
    step0 = m.bnn._step
    got = m.bnn.predict_moments_host(x, 6, first_fraction=0.2, n_chunks=3)  # 600 + 1200 + 1200 windows
    m.bnn._step = step0  # same Philox key again
    ref = m.bnn.predict_moments(x.to(DEV), 6)
    for a, b in zip(got, ref):
        np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-5, atol=1e-6)


def test_predict_task_writer_and_deep_ensemble(tmp_path):
    """The predict task end to end (tasks/predict.py:52-64 -> results/predictions.py:31-56): HNN members write
    `<method>_<run>_<subset>.parquet` through predict_step, the frames are read back and mixed by deep_ensemble."""
    import pandas as pd
    from bayesrul_b200.compat import HNN, Inception, deep_ensemble, write_predictions
    g = torch.Generator().manual_seed(3)
    batches = [(torch.randn(n, 30, 18, generator=g), torch.rand(n, generator=g) * 100) for n in (40, 40, 17)]
    frames = []
    for run in range(3):
        torch.manual_seed(100 + run)
        m = HNN(Inception(30, 18), None, mc_samples=1, p_dropout=0, device=DEV)  # HNN applies weights_init itself
        name = f"HNN_{run:03d}_test.parquet"
        f = write_predictions(m, batches, tmp_path, name)
        assert list(f.columns) == ["labels", "preds", "stds"] and len(f) == 97  # no MC-dropout: no variance split (frequentist.py:132-151)
        np.testing.assert_allclose(f.labels.values, torch.cat([b[1] for b in batches]).numpy(), rtol=1e-6)
        frames.append(pd.read_parquet(tmp_path / name).assign(model=f"HNN_{run:03d}", method="HNN"))
    df = pd.concat(frames).reset_index(drop=True)
    for ens in ((0, 1), (0, 2), (0, 1, 2)):
        o = deep_ensemble(df[df.model.isin([f"HNN_{k:03d}" for k in ens])])
        mu = np.stack([frames[k].preds.values for k in ens]).astype(np.float64)
        sd = np.stack([frames[k].stds.values for k in ens]).astype(np.float64)
        assert list(o.columns) == ["preds", "labels", "stds"] and len(o) == 97
        np.testing.assert_allclose(o.labels.values, frames[0].labels.values)
        np.testing.assert_allclose(o.preds.values, mu.mean(0), rtol=1e-5)
        np.testing.assert_allclose(o.stds.values, np.sqrt((mu**2 + sd**2).mean(0) - mu.mean(0) ** 2), rtol=1e-3, atol=1e-4)
