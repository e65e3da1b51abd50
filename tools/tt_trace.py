"""In-kernel clock64 timeline of CTA (0, layer) of the level-fused training kernels (brl_tt_trace): where a CTA's time goes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayesrul_b200 import Engine, Noise, _lib
from bayesrul_b200.compat.nets import init_flat_params

mode = sys.argv[1] if len(sys.argv) > 1 else "lrt"
dev = torch.device("cuda", 0)
e = Engine("inception", dev)
e.set_gemm_backend("fused")
e.set_step_graph(False)
B = 256
g = torch.Generator().manual_seed(1)
x = torch.randn(B, 30, 18, generator=g).to(dev); y = (torch.rand(B, generator=g) * 100).to(dev)
mu = init_flat_params("inception", 12345).to(dev); sg = torch.full_like(mu, 1.351e-3)
particles = 2 if mode == "flipout" else 1
kw = dict(mode=mode, guide="normal", particles=particles, prior_loc=0.0, prior_scale=0.138793, dataset_size=238150)
for i in range(3):
    e.elbo_step(x, y, mu, sg, noise=Noise(seed=i), **kw)
buf = torch.zeros(6 * 4 * 16, dtype=torch.int64, device=dev)
_lib.check(e.lib.brl_tt_trace(e.ctx, buf.data_ptr()))
e.elbo_step(x, y, mu, sg, noise=Noise(seed=9), **kw)
torch.cuda.synchronize()
_lib.check(e.lib.brl_tt_trace(e.ctx, None))
t = buf.cpu().view(6, 4, 16)
names = ["fwd A", "fwd B", "fwd C", "bwd {b2b,b3b,b1,b4}", "bwd {b2a,b3a}", "bwd module 1"]
lab = ["setup", "issue copies", "copies land", "operands built", "MMAs done", "epilogue", "dW flush", "end sync"]
for k in range(6):
    for l in range(4):
        s = t[k, l]
        if s[0] == 0:
            continue
        d = [(int(s[i + 1]) - int(s[i])) for i in range(7) if s[i + 1] > 0]
        extra = ""
        if s[8] > 0:
            extra = f" || prologue={int(s[8]) - int(s[0])} out_grad={int(s[9]) - int(s[8])} d1+sts={int(s[10]) - int(s[9])} sums={int(s[1]) - int(s[10])}"
        print(f"{names[k]:22s} layer-slot {l}: total {int(max(s[:8])) - int(s[0]):7d} cyc | " + " ".join(f"{n}={v}" for n, v in zip(lab[1:], d)) + extra)
