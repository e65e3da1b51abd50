"""Linear (default) or Conv-D3 net (argv[1] = conv), weight-sampling predict (B = 10 000 x S = 100, q_scale 1.351e-3): tcgen05 engine vs the per-layer engines."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesrul_b200 import Engine, Noise
from bayesrul_b200.compat.nets import init_flat_params

dev = torch.device("cuda", 0)
net = sys.argv[1] if len(sys.argv) > 1 else "linear"
e = Engine(net, dev)
B, S = 10000, 100
x = torch.randn(B, 30, 18, device=dev)
mu = init_flat_params(net, 12345).to(dev)
sg = torch.full_like(mu, 1.351e-3)
for engine, be in (("tc", "simt"), ("simt", "simt"), ("simt", "tc")):
    e.set_gemm_backend(be)
    for i in range(2):
        e.predict_moments(x, mu, sg, S=S, guide="normal", noise=Noise(seed=i), engine=engine)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    n = 3
    for i in range(n):
        e.predict_moments(x, mu, sg, S=S, guide="normal", noise=Noise(seed=10 + i), engine=engine)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / n
    print(f"engine {engine:4s} per-layer backend {be:4s}: {ms:8.3f} ms/step = {B * S / ms / 1e3:8.1f} M window-samples/s; status {e.tc_status()}")
