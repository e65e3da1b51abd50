"""Probe: LRT / Flipout ELBO step time vs minibatch size (graph replay, fp32 back-end)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesrul_b200 import Engine, Noise
from bayesrul_b200.compat.nets import init_flat_params

dev = torch.device("cuda:0")
e = Engine("inception", dev)
mu = init_flat_params("inception", 12345).to(dev)
for mode, parts, q, ps in (("lrt", 1, 1.351e-3, 0.138793), ("flipout", 2, 2.14e-4, 0.198768)):
    sg = torch.full_like(mu, q)
    for B in (64, 128, 256, 512, 1024, 4096):
        g = torch.Generator().manual_seed(B)
        x = torch.randn(B, 30, 18, generator=g).to(dev)
        y = (torch.rand(B, generator=g) * 100).to(dev)
        for i in range(6):
            e.elbo_step(x, y, mu, sg, mode=mode, particles=parts, prior_scale=ps, dataset_size=238150, noise=Noise(seed=i))
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(50):
            e.elbo_step(x, y, mu, sg, mode=mode, particles=parts, prior_scale=ps, dataset_size=238150, noise=Noise(seed=i))
        b.record()
        torch.cuda.synchronize()
        t = a.elapsed_time(b) / 50
        print(f"{mode} B={B}: {t:.3f} ms/step = {B / t:.0f} k windows/s", flush=True)
