"""ELBO step time (graph replay + ClippedAdam, B = 256) of the fp32 FFMA back-end vs the level-fused tcgen05 back-end."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesrul_b200 import Engine, Noise
from bayesrul_b200.compat.nets import init_flat_params

dev = torch.device("cuda", 0)
eng = Engine("inception", dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator().manual_seed(1)
x = torch.randn(B, 30, 18, generator=g).to(dev)
y = (torch.rand(B, generator=g) * 100).to(dev)
mu0 = init_flat_params("inception", 12345).to(dev)
for mode, particles, q, ps in (("lrt", 1, 1.351e-3, 0.138793), ("flipout", 2, 2.14e-4, 0.198768)):
    for backend in ("simt", "fused"):
        eng.set_gemm_backend(backend)
        mu = mu0.clone(); ls = torch.full_like(mu, float(torch.log(torch.tensor(q)))); sg = torch.full_like(mu, q)
        opt = [torch.zeros_like(mu) for _ in range(4)]
        it = [0]
        def step():
            it[0] += 1
            r = eng.elbo_step(x, y, mu, sg, mode=mode, guide="normal", particles=particles, prior_loc=0.0, prior_scale=ps,
                              dataset_size=238150, noise=Noise(seed=5000 + it[0]))
            eng.clipped_adam_vi(mu, ls, sg, r["grad_mu"], r["grad_log_sigma"], *opt, it[0], 1e-3, (0.95, 0.999), 1e-8, 15.0)
            return r
        for _ in range(10): r = step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 200
        a.record()
        for _ in range(n): r = step()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        print(f"{mode:8s} {backend:6s} B={B}: {ms:.4f} ms/step  {B / ms:.0f} k windows/s  loss {r['scalars'][0].item():.5f}  status {eng.tc_status()}")
eng.set_gemm_backend("simt")
