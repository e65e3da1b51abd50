"""Debug: timeline (clock64) of CTA 0 of tc_conv_kernel -- issuer vs two epilogue warps, first 16 items."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesrul_b200 import Engine, Noise
from bayesrul_b200.compat.nets import init_flat_params as init_params

dev = "cuda:0"
e = Engine("inception", dev)
g = torch.Generator().manual_seed(0)
x = torch.randn(10000, 30, 18, generator=g).to(dev)
mu = init_params("inception", 1).to(dev)
sg = torch.full_like(mu, 1.351e-3)
for _ in range(2):
    e.predict_moments(x, mu, sg, S=16, guide="normal", noise=Noise(seed=1), engine="tc")
buf = torch.zeros(16 * 128, dtype=torch.int64, device=dev)
e.lib.brl_tc_trace(e.ctx, C.c_void_p(buf.data_ptr()))
e.predict_moments(x, mu, sg, S=16, guide="normal", noise=Noise(seed=1), engine="tc")
torch.cuda.synchronize()
e.lib.brl_tc_trace(e.ctx, None)
t = buf.cpu().view(16, 128)
t0 = int(t[2, 0])
names = {}
for w, base in (("w0", 0), ("w15", 18)):
    for pi, ph in enumerate("ABC"):
        for k in range(2):
            names[base + pi * 6 + 3 * k] = f"{w} {ph}{k} wait>"
            names[base + pi * 6 + 3 * k + 1] = f"{w} {ph}{k} <wait"
            names[base + pi * 6 + 3 * k + 2] = f"{w} {ph}{k} end"
for k in range(2):
    names[40 + 2 * k] = f"iss B{k} ready"
    names[41 + 2 * k] = f"iss B{k} issued"
    names[44 + 2 * k] = f"iss C{k} ready"
    names[45 + 2 * k] = f"iss C{k} issued"
    names[48 + k] = f"iss A'{k} issued"
    names[50 + k] = f"iss X'{k} load issued"
    names[52 + k] = f"iss A'{k} xfull seen"
for it in range(2, 6):
    ev = sorted((int(t[it, s]) - t0, names[s]) for s in names if int(t[it, s]) != 0)
    print(f"--- item {it}")
    for ph, nm in ((0, "A0"), (1, "B0")):
        print(f"   arrive {nm} per warp:", " ".join(str(int(t[it, 64 + ph * 16 + w]) - t0) for w in range(16)))
    prev = None
    for c, n in ev:
        print(f"{c:8d}  {n}")
