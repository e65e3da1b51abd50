import sys, math, torch
sys.path.insert(0, '/root/repo')
from bayesrul_b200 import Engine, Noise
from bayesrul_b200.compat.nets import init_flat_params


def synth(net, B, seed=0, sigma=0.05):  # (the oracle is test infrastructure: tools build their inputs from the package)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 30, 18, generator=g)
    y = torch.rand(B, generator=g) * 100
    mu = init_flat_params(net, seed + 1)
    return x, y, mu, torch.full_like(mu, sigma)


DEV = "cuda:0"
for net in ("linear", "inception"):
    e = Engine(net, DEV)
    for B in (96, 256):
        x, y, mu, sg = synth(net, B, seed=21, sigma=0.05)
        x, y, mu, sg = x.to(DEV), y.to(DEV), mu.to(DEV), sg.to(DEV)
        for S in (1, 2, 3):
            w = e.sample_weights(mu, sg, "normal", S, Noise(seed=5))
            for mode in ("flipout", "lrt", "ws"):
                def fn():
                    if mode == "flipout":
                        return e.forward(x, "flipout", theta=mu, wsamp=w, S=S, noise=Noise(seed=11)).clone()
                    if mode == "lrt":
                        return e.forward(x, "lrt", theta=mu, sigma=sg, S=S, noise=Noise(seed=11)).clone()
                    return e.forward(x, "ws", wsamp=w, S=S).clone()
                e.set_gemm_backend(0); a = fn()
                e.set_gemm_backend(7); b = fn()
                e.set_gemm_backend(0)
                d = (a - b).abs().amax(dim=(1, 2))
                print(net, "B", B, "S", S, mode, "per-sample max diff", [f"{v:.2e}" for v in d.tolist()], "scale", f"{a.abs().max().item():.2e}")
