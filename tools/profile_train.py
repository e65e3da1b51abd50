"""Small driver for ncu: a few ELBO train steps (LRT, B=256) through the C ABI."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesrul_b200 import Engine, Noise
from bayesrul_b200.compat.nets import init_flat_params as init_params

mode = sys.argv[1] if len(sys.argv) > 1 else "lrt"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
backend = sys.argv[3] if len(sys.argv) > 3 else "simt"
dev = "cuda:0"
e = Engine("inception", dev)
e.set_gemm_backend(backend)
particles = 2 if mode == "flipout" else 1
g = torch.Generator().manual_seed(0)
x = torch.randn(256, 30, 18, generator=g).to(dev)
y = (torch.rand(256, generator=g) * 100).to(dev)
mu = init_params("inception", 1).to(dev)
sg = torch.full_like(mu, 1.351e-3)
for i in range(steps):
    r = e.elbo_step(x, y, mu, sg, mode=mode, guide="normal", particles=particles, prior_loc=0.0, prior_scale=0.138793,
                    dataset_size=238150, noise=Noise(seed=i))
torch.cuda.synchronize()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for i in range(20):
    r = e.elbo_step(x, y, mu, sg, mode=mode, guide="normal", particles=particles, prior_loc=0.0, prior_scale=0.138793,
                    dataset_size=238150, noise=Noise(seed=100 + i))
t1.record(); torch.cuda.synchronize()
print("ms/step", t0.elapsed_time(t1) / 20, "loss", r["scalars"][0].item())

# CUDA-graph replay of the same step (fixed seed): GPU time without host launch overhead
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for i in range(2):
        e.elbo_step(x, y, mu, sg, mode=mode, guide="normal", particles=particles, prior_loc=0.0, prior_scale=0.138793,
                    dataset_size=238150, noise=Noise(seed=7))
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        r = e.elbo_step(x, y, mu, sg, mode=mode, guide="normal", particles=particles, prior_loc=0.0, prior_scale=0.138793,
                        dataset_size=238150, noise=Noise(seed=7))
torch.cuda.synchronize()
g.replay(); torch.cuda.synchronize()
t0.record()
for i in range(50):
    g.replay()
t1.record(); torch.cuda.synchronize()
print("graph ms/step", t0.elapsed_time(t1) / 50, "loss", r["scalars"][0].item())
