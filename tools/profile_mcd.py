"""Small driver for ncu: MC-dropout predictive passes (configs[1]) on the tcgen05 engine."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesrul_b200 import Engine, Noise
from bayesrul_b200.compat.nets import init_flat_params

dev = torch.device("cuda:0")
e = Engine("inception", dev)
mu = init_flat_params("inception", 12345).to(dev)
g = torch.Generator().manual_seed(0)
x = torch.randn(10000, 30, 18, generator=g).to(dev)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for _ in range(3):
    e.predict_moments(x, mu, None, S=S, guide=None, p_dropout=0.241437, noise=Noise(seed=4048), engine="tc")
torch.cuda.synchronize()
e.tc_timing(True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    e.predict_moments(x, mu, None, S=S, guide=None, p_dropout=0.241437, noise=Noise(seed=4048), engine="tc")
b.record()
torch.cuda.synchronize()
kt = e.tc_timing_read()
print(f"mcd step {a.elapsed_time(b) / 3:.3f} ms; conv {kt['tc_conv_kernel'][0] / kt['tc_conv_kernel'][1]:.4f} ms/launch x{kt['tc_conv_kernel'][1] // 3}, "
      f"fc {kt['tc_fc_kernel'][0] / kt['tc_fc_kernel'][1]:.4f} ms/launch", flush=True)
