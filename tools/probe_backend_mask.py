"""Probe: ELBO step time for every GEMM back-end bit mask (1 forward, 2 input-gradient, 4 weight-gradient kernels on tcgen05 TF32)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesrul_b200 import Engine, Noise, _lib
from bayesrul_b200.compat.nets import init_flat_params

dev = torch.device("cuda:0")
e = Engine("inception", dev)
mu = init_flat_params("inception", 12345).to(dev)
g = torch.Generator().manual_seed(0)
x = torch.randn(256, 30, 18, generator=g).to(dev)
y = (torch.rand(256, generator=g) * 100).to(dev)
for mode, parts, q, ps in (("lrt", 1, 1.351e-3, 0.138793), ("flipout", 2, 2.14e-4, 0.198768)):
    sg = torch.full_like(mu, q)
    for mask in range(8):
        _lib.check(e.lib.brl_set_gemm_backend(e.ctx, mask))
        for i in range(10):
            e.elbo_step(x, y, mu, sg, mode=mode, particles=parts, prior_scale=ps, dataset_size=238150, noise=Noise(seed=i))
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(100):
            e.elbo_step(x, y, mu, sg, mode=mode, particles=parts, prior_scale=ps, dataset_size=238150, noise=Noise(seed=i))
        b.record()
        torch.cuda.synchronize()
        print(f"{mode} mask {mask}: {a.elapsed_time(b) / 100:.3f} ms/step  status {e.gemm_status()}", flush=True)
