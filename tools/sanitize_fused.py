"""Smallest eager case of the level-fused back-end (one LRT, Flipout x2 particles, weight-sampling and HNN step), the command to
put under a memory checker:  compute-sanitizer --tool memcheck python tools/sanitize_fused.py 9
(compute-sanitizer is closed on the round-2 GPU pool, so only the plain run -- status 0 -- has been exercised there.)"""
import os, sys
os.environ["BRL_NO_GRAPH"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesrul_b200 import Engine, Noise
from bayesrul_b200.compat.nets import init_flat_params

dev = torch.device("cuda", 0)
e = Engine("inception", dev)
e.set_gemm_backend("fused")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 9
g = torch.Generator().manual_seed(1)
x = torch.randn(B, 30, 18, generator=g).to(dev); y = (torch.rand(B, generator=g) * 100).to(dev)
mu = init_flat_params("inception", 12345).to(dev); sg = torch.full_like(mu, 0.01)
kw = dict(prior_loc=0.0, prior_scale=0.138793, dataset_size=238150)
for mode, guide, particles in (("lrt", "normal", 1), ("flipout", "normal", 2), ("ws", "radial", 1)):
    r = e.elbo_step(x, y, mu, sg, mode=mode, guide=guide, particles=particles, noise=Noise(seed=3), **kw)
    print(mode, guide, float(r["scalars"][0]), float(r["grad_mu"].abs().sum()))
r = e.hnn_step(x, y, mu, 0.241437, Noise(seed=4))
print("hnn", float(r["scalars"][0]), float(r["grad"].abs().sum()))
torch.cuda.synchronize()
print("status", e.tc_status())
