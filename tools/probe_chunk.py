"""Probe: predictive step time vs the number of MC samples per fused launch (BRL_TC_SC + chunk=)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesrul_b200 import Engine, Noise
from bayesrul_b200.compat.nets import init_flat_params

dev = torch.device("cuda:0")
e = Engine("inception", dev)
mu = init_flat_params("inception", 12345).to(dev)
sg = torch.full_like(mu, 1.351e-3)
g = torch.Generator().manual_seed(0)
x = torch.randn(10000, 30, 18, generator=g).to(dev)
sc = int(os.environ.get("BRL_TC_SC", "32"))
flush = torch.zeros(64 * 1024 * 1024, device=dev)
for _ in range(3):
    e.predict_moments(x, mu, sg, S=100, guide="normal", noise=Noise(seed=1), engine="tc", chunk=sc)
torch.cuda.synchronize()
ts = []
for _ in range(6):
    flush.add_(1.0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    e.predict_moments(x, mu, sg, S=100, guide="normal", noise=Noise(seed=1), engine="tc", chunk=sc)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(f"samples per launch {sc}: {sum(ts) / len(ts):.3f} ms/step", flush=True)
