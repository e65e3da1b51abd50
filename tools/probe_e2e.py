"""Probe: BNN.predict_step (pinned host batch -> numpy) for the host-chunking settings in BRL_HOST_CHUNKS / BRL_HOST_FIRST."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesrul_b200.compat import BNN, Inception

dev = torch.device("cuda:0")
m = BNN(Inception(30, 18), None, 0, 1, 100, 238150, "lrt", 0.0, 0.138793, "normal", 1.351e-3, device=dev, engine="tc")
m.on_predict_start()
g = torch.Generator().manual_seed(0)
hx = [(torch.randn(10000, 30, 18, generator=g).pin_memory(), torch.rand(10000, generator=g).pin_memory()) for _ in range(2)]
flush = torch.zeros(64 * 1024 * 1024, device=dev)
for i in range(3):
    m.predict_step(hx[i % 2], i)
torch.cuda.synchronize()
ts = []
for i in range(6):
    flush.add_(1.0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    m.predict_step(hx[i % 2], i)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(f"chunks={os.environ.get('BRL_HOST_CHUNKS')} first={os.environ.get('BRL_HOST_FIRST')}: {sum(ts) / len(ts):.3f} ms/step "
      f"({1e6 / (sum(ts) / len(ts)) / 1e3:.1f} M window-samples/s)", flush=True)
