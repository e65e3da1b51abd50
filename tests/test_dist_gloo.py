"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard ranges, Chan moment merge across
sample shards, ensemble mixture across ranks, flat gradient all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bnn_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bayesrul_b200 import dist as D

    g = torch.Generator().manual_seed(0)
    S, B = 11, 50
    out = torch.rand(S, B, 2, generator=g) * 5 + 0.1
    want = O.predictive_moments(out)
    # (1) MC samples sharded over ranks (uneven: 6 + 5)
    a, b = D.shard_range(S, rank, world)
    loc = O.predictive_moments(out[a:b])
    got = D.all_gather_moments(b - a, loc[0], loc[2], loc[3])
    ok = all(torch.allclose(u, v, rtol=1e-5, atol=1e-6) for u, v in zip(got, want))
    # (2) windows sharded over ranks: concatenation in rank order
    a, b = D.shard_range(B, rank, world)
    cat = D.gather_predictions([w[a:b].contiguous() for w in want])
    ok = ok and all(torch.equal(u, v) for u, v in zip(cat, want))
    # (3) ensemble members over ranks
    mu_m, sd_m = torch.rand(5, 40, generator=g) * 50, torch.rand(5, 40, generator=g) + 0.5
    a, b = D.shard_range(5, rank, world)
    mu, sd = D.mixture_across_ranks(mu_m[a:b], sd_m[a:b])
    wm, ws = O.deep_ensemble_moments(mu_m, sd_m)
    ok = ok and torch.allclose(mu, wm, rtol=1e-5) and torch.allclose(sd, ws, rtol=1e-4)
    # (4) data-parallel ELBO gradients: mean over ranks of the per-rank (plate N/B_r) gradients == global batch
    net = "conv"
    x = torch.randn(8, 30, 18, generator=g, dtype=torch.float64)
    y = torch.rand(8, generator=g, dtype=torch.float64) * 50
    mu0 = O.init_params(net, 1, torch.float64)
    sg0 = torch.full_like(mu0, 0.05)
    nz = O.make_injected_noise(net, 8, "lrt", g, dtype=torch.float64)
    kw = dict(mode="lrt", guide="normal", prior_loc=0.0, prior_scale=0.2, dataset_size=1000)
    full = O.elbo_loss_and_grads(net, x, y, mu0, sg0, noises=[O.InjectedNoise(nz)], **kw)
    a, b = D.shard_range(8, rank, world)
    nzr = {k: v[a:b] for k, v in nz.items()}
    part = O.elbo_loss_and_grads(net, x[a:b], y[a:b], mu0, sg0, noises=[O.InjectedNoise(nzr)], **kw)
    part["scalars"] = torch.stack([part["loss"], part["nll_sum"], part["kl"], torch.zeros(())])
    P = mu0.numel()  # the engine's flat result buffer: [grad_mu | grad_log_sigma | 4 scalars | grad_sigma] (Engine.elbo_step)
    flat = torch.cat([part["grad_mu"], part["grad_log_sigma"], torch.zeros(4, dtype=torch.float64), part["grad_sigma"]])
    part.update(flat=flat, grad_mu=flat[:P], grad_log_sigma=flat[P:2 * P], grad_sigma=flat[2 * P + 4:])
    red = D.allreduce_elbo_grads(part)
    ok = ok and torch.allclose(red["grad_mu"], full["grad_mu"], rtol=1e-9, atol=1e-14)
    ok = ok and torch.allclose(red["grad_log_sigma"], full["grad_log_sigma"], rtol=1e-9, atol=1e-14)
    ok = ok and abs(red["scalars"][0].item() - full["loss"].item()) < 1e-12
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_shard_range():
    from bayesrul_b200.dist import shard_range
    for n, w in ((10, 3), (7, 8), (1000000, 8), (0, 2)):
        blocks = [shard_range(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(180)
def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(30)
    assert sorted(res) == [(0, True), (1, True)]
