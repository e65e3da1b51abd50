"""How many SMs would an L2-fed `tc_fc_kernel` need?  (conv + fc co-resident in one launch.)

Runs the predictive step at a size whose feature tensor stays L2-resident (B x S window-samples x 4800 B <= ~60 MB) and at the
bench size (4.8 GB: HBM-fed), with the fc grid restricted through the BRL_FC_GRID knob (read once per process: run this script once
per grid size), and prints the fc kernel's time per 128-window tile per CTA and the implied per-SM feed rate.

usage: BRL_FC_GRID=16 python tools/probe_fc_l2.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesrul_b200 import Engine, Noise
from bayesrul_b200.compat.nets import init_flat_params

dev = torch.device("cuda", 0)
e = Engine("inception", dev)
grid = int(os.environ.get("BRL_FC_GRID", "148"))
mu = init_flat_params("inception", 12345).to(dev)
sg = torch.full_like(mu, 1.351e-3)
for B, S, tag in ((1024, 12, "L2-resident (59 MB)"), (10000, 100, "HBM (4.8 GB)")):
    x = torch.randn(B, 30, 18, device=dev)
    for _ in range(3):
        e.predict_moments(x, mu, sg, S=S, guide="normal", noise=Noise(seed=1), engine="tc")
    torch.cuda.synchronize()
    e.tc_timing(True)
    n = 10
    for i in range(n):
        e.predict_moments(x, mu, sg, S=S, guide="normal", noise=Noise(seed=2 + i), engine="tc")
    torch.cuda.synchronize()
    tm = e.tc_timing_read()
    e.tc_timing(False)
    ms = (tm["tc_conv_kernel"][0], tm["tc_fc_kernel"][0])
    tiles = S * ((B + 127) // 128)
    g = min(grid, tiles)
    conv_ms, fc_ms = ms[0] / n, ms[1] / n
    per_tile_us = fc_ms * 1e3 / (tiles / g)
    feed = (614400 + 38 * 8192) / per_tile_us / 1e3  # GB/s per SM: feature tile + fc weight k-blocks
    print(f"fc grid {g:3d}  {tag:22s} B={B} S={S}: conv {conv_ms:.3f} ms  fc {fc_ms:.3f} ms  = {per_tile_us:.2f} us per tile per CTA "
          f"({feed:.0f} GB/s per SM); fc SMs needed to keep pace with the conv kernel: {g * fc_ms / conv_ms:.1f}")
print("status", e.tc_status())
