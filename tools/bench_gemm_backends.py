"""fp32 SIMT vs tcgen05 TF32 per-layer kernels at predict-sized row counts (B = 10000 windows, S weight samples)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesrul_b200 import Engine, Noise
from bayesrul_b200.compat.nets import init_flat_params as init_params

dev = "cuda:0"
B, S = 10000, 8
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 30, 18, generator=g).to(dev)
for net in ("inception", "conv", "linear"):
    e = Engine(net, dev)
    mu = init_params(net, 1).to(dev)
    sg = torch.full_like(mu, 1.351e-3)
    for backend in ("simt", "tc"):
        e.set_gemm_backend(backend)
        for _ in range(2):
            e.predict_moments(x, mu, sg, S=S, guide="normal", noise=Noise(seed=1), engine="simt")
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(3):
            r = e.predict_moments(x, mu, sg, S=S, guide="normal", noise=Noise(seed=1), engine="simt")
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 3
        print(f"{net:10s} {backend:5s} {ms:8.2f} ms  {B * S / ms / 1e3:8.2f} M window-samples/s  status {e.gemm_status()}")
    e.set_gemm_backend("simt")
