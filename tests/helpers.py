"""Shared helpers of the GPU parity tests (oracle <-> engine noise conversion)."""
import numpy as np
import torch

from oracle import bnn_oracle as O


def injected_to_engine(net, noises, B, device):
    """list (one per MC sample) of oracle noise dicts -> bayesrul_b200.Noise with [S,...] tensors."""
    from bayesrul_b200 import Noise

    S = len(noises)
    nz = Noise()

    def stack(key):
        return torch.stack([n[key].float() for n in noises]).contiguous().to(device)

    keys = noises[0].keys()
    if "weight_eps" in keys:
        nz.weight_eps = stack("weight_eps")
    if "radial_r" in keys:
        nz.radial_r = stack("radial_r")
    for k in keys:
        if "." in k:
            name, layer = k.split(".")
            getattr(nz, name)[int(layer)] = stack(k).reshape(S, B, -1).contiguous()
    return nz


def assert_close(a, b, rtol=1e-3, atol_scale=1e-5, what=""):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    scale = b.abs().max().item() + 1e-30
    err = (a - b).abs()
    tol = rtol * b.abs() + atol_scale * scale
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} mismatches, max err {err.max().item():.3e}, scale {scale:.3e}"


def synth(net, B, seed=0, sigma=0.05, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 30, 18, generator=g, dtype=torch.float64).to(dtype)
    y = (torch.rand(B, generator=g, dtype=torch.float64) * 100).to(dtype)
    mu = O.init_params(net, seed + 1, dtype)
    sg = torch.full_like(mu, sigma)
    return x, y, mu, sg
