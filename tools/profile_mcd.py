"""Small driver for ncu: the MC-dropout predictive step of configs[1] (Inception, p = 0.241437, B = 10 000 x S = 100 fused Philox masks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesrul_b200 import Engine, Noise
from bayesrul_b200.compat.nets import init_flat_params

dev = torch.device("cuda", 0)
e = Engine("inception", dev)
B, S = int(sys.argv[1]) if len(sys.argv) > 1 else 10000, int(sys.argv[2]) if len(sys.argv) > 2 else 100
x = torch.randn(B, 30, 18, device=dev)
mu = init_flat_params("inception", 12345).to(dev)
for i in range(3):
    e.predict_moments(x, mu, None, S=S, guide=None, p_dropout=0.241437, noise=Noise(seed=i), engine="tc")
torch.cuda.synchronize()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for i in range(3):
    e.predict_moments(x, mu, None, S=S, guide=None, p_dropout=0.241437, noise=Noise(seed=10 + i), engine="tc")
t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / 3
print(f"mcd predict B={B} S={S}: {ms:.3f} ms/step = {B * S / ms / 1e3:.1f} M window-samples/s; status {e.tc_status()}")
