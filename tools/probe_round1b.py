"""Probe: host enqueue time vs GPU time of one ELBO step; predict step with reduced conv / fc grids (env knobs)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesrul_b200 import Engine, Noise
from bayesrul_b200.compat.nets import init_flat_params

dev = torch.device("cuda:0")
e = Engine("inception", dev)
mu = init_flat_params("inception", 12345).to(dev)
what = sys.argv[1] if len(sys.argv) > 1 else "train"
if what == "train":
    g = torch.Generator().manual_seed(0)
    x = torch.randn(256, 30, 18, generator=g).to(dev)
    y = (torch.rand(256, generator=g) * 100).to(dev)
    for mode, parts, q, ps in (("lrt", 1, 1.351e-3, 0.138793), ("flipout", 2, 2.14e-4, 0.198768)):
        sg = torch.full_like(mu, q)
        for backend in ("simt", "tc"):
            e.set_gemm_backend(backend)
            for i in range(10):
                e.elbo_step(x, y, mu, sg, mode=mode, particles=parts, prior_scale=ps, dataset_size=238150, noise=Noise(seed=i))
            torch.cuda.synchronize()
            n = 100
            t0 = time.perf_counter()
            for i in range(n):
                e.elbo_step(x, y, mu, sg, mode=mode, particles=parts, prior_scale=ps, dataset_size=238150, noise=Noise(seed=i))
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            print(f"{mode} {backend}: host enqueue {1e3 * (t1 - t0) / n:.3f} ms/step, total {1e3 * (t2 - t0) / n:.3f} ms/step", flush=True)
else:
    g = torch.Generator().manual_seed(0)
    x = torch.randn(10000, 30, 18, generator=g).to(dev)
    sg = torch.full_like(mu, 1.351e-3)
    for _ in range(3):
        e.predict_moments(x, mu, sg, S=100, guide="normal", noise=Noise(seed=1), engine="tc")
    torch.cuda.synchronize()
    e.tc_timing(True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        e.predict_moments(x, mu, sg, S=100, guide="normal", noise=Noise(seed=1), engine="tc")
    b.record()
    torch.cuda.synchronize()
    kt = e.tc_timing_read()
    print(f"CONV_GRID={os.environ.get('BRL_CONV_GRID')} FC_GRID={os.environ.get('BRL_FC_GRID')}: step {a.elapsed_time(b) / 5:.3f} ms; "
          f"conv {kt['tc_conv_kernel'][0] / kt['tc_conv_kernel'][1]:.4f} ms/launch, fc {kt['tc_fc_kernel'][0] / kt['tc_fc_kernel'][1]:.4f} ms/launch", flush=True)
