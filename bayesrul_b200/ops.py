"""`torch.ops.bayesrul_b200.*`: the tensor-level operator set SURVEY.md section 8(b) names, registered with torch.library on top of
the C ABI (include/bayesrul_b200.h) so that the path composes with autograd / torch.compile-free graphs like any other op:

    forward(x, theta, sigma, wsamp, net, mode, S, p_dropout, seed, engine)            -> out [S,B,2]
    predict_moments(x, mu, sigma, net, S, guide, p_dropout, seed, engine)             -> pred, std, ep_var, al_var  [B] each
    elbo_step(x, y, mu, sigma, net, mode, guide, particles, prior_loc, prior_scale,
              dataset_size, seed, backend)                                            -> loss_and_scalars f64[4], grad_mu, grad_log_sigma, out
    mixture_moments(mu_m, sigma_m)                                                     -> mu, sigma  (deepens.py:21-24)
    elbo_loss(...)  = `elbo_step` as a differentiable scalar: `loss.backward()` fills mu.grad and log_sigma.grad with the
                      gradients the fused step computed (torch.autograd.Function)

`net` / `mode` / `guide` / `engine` / `backend` are the strings of bayesrul_b200.engine.  Engines (one brl_ctx per net and device)
are cached.  Noise is native Philox keyed by `seed`; injected-noise parity goes through Engine directly.  There is no CPU
implementation: the ops raise on non-CUDA tensors.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from .engine import Engine, Noise

_ENGINES: Dict[Tuple[str, int], Engine] = {}


def _engine(net: str, t: torch.Tensor) -> Engine:
    if not t.is_cuda:
        raise RuntimeError("bayesrul_b200 ops need CUDA tensors (B200 / sm_100a, no CPU fallback)")
    key = (net, t.device.index)
    if key not in _ENGINES:
        _ENGINES[key] = Engine(net, t.device)
    return _ENGINES[key]


def _opt(t: torch.Tensor):
    return t if t.numel() > 0 else None


@torch.library.custom_op("bayesrul_b200::forward", mutates_args=())
def forward(x: torch.Tensor, theta: torch.Tensor, sigma: torch.Tensor, wsamp: torch.Tensor, net: str, mode: str, S: int,
            p_dropout: float, seed: int, engine: str) -> torch.Tensor:
    """Pass empty tensors for the operands a mode does not use (theta: det/lrt/flipout, sigma: lrt, wsamp: ws/flipout)."""
    return _engine(net, x).forward(x, mode, theta=_opt(theta), sigma=_opt(sigma), wsamp=_opt(wsamp), S=S, p_dropout=p_dropout,
                                   noise=Noise(seed=seed), engine=engine)


@forward.register_fake
def _(x, theta, sigma, wsamp, net, mode, S, p_dropout, seed, engine):
    return x.new_empty(S, x.shape[0], 2)


@torch.library.custom_op("bayesrul_b200::predict_moments", mutates_args=())
def predict_moments(x: torch.Tensor, mu: torch.Tensor, sigma: torch.Tensor, net: str, S: int, guide: str, p_dropout: float, seed: int,
                    engine: str) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """guide "normal" | "radial" | "none" (MC-dropout / deterministic weights `mu`)."""
    g = None if guide == "none" else guide
    return _engine(net, x).predict_moments(x, mu, _opt(sigma), S=S, guide=g, p_dropout=p_dropout, noise=Noise(seed=seed), engine=engine)


@predict_moments.register_fake
def _(x, mu, sigma, net, S, guide, p_dropout, seed, engine):
    return tuple(x.new_empty(x.shape[0]) for _ in range(4))


@torch.library.custom_op("bayesrul_b200::elbo_step", mutates_args=())
def elbo_step(x: torch.Tensor, y: torch.Tensor, mu: torch.Tensor, sigma: torch.Tensor, net: str, mode: str, guide: str,
              particles: int, prior_loc: float, prior_scale: float, dataset_size: int, seed: int,
              backend: str) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    e = _engine(net, x)
    e.set_gemm_backend(backend)
    r = e.elbo_step(x, y, mu, sigma, mode=mode, guide=guide, particles=particles, prior_loc=prior_loc, prior_scale=prior_scale,
                    dataset_size=dataset_size, noise=Noise(seed=seed))
    return r["scalars"], r["grad_mu"].clone(), r["grad_log_sigma"].clone(), r["out"]


@elbo_step.register_fake
def _(x, y, mu, sigma, net, mode, guide, particles, prior_loc, prior_scale, dataset_size, seed, backend):
    return (x.new_empty(4, dtype=torch.float64), torch.empty_like(mu), torch.empty_like(mu), x.new_empty(particles, x.shape[0], 2))


@torch.library.custom_op("bayesrul_b200::mixture_moments", mutates_args=())
def mixture_moments(mu_m: torch.Tensor, sigma_m: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    return _engine("inception", mu_m).mixture_moments(mu_m, sigma_m)


@mixture_moments.register_fake
def _(mu_m, sigma_m):
    return mu_m.new_empty(mu_m.shape[1]), mu_m.new_empty(mu_m.shape[1])


class _ElboLoss(torch.autograd.Function):
    """ELBO of one minibatch as a differentiable scalar of (mu, log_sigma): forward runs the fused step (which already computes
    the gradients), backward hands them out scaled by the incoming gradient."""

    @staticmethod
    def forward(ctx, mu, log_sigma, x, y, net, mode, guide, particles, prior_loc, prior_scale, dataset_size, seed, backend):
        sc, g_mu, g_ls, _ = torch.ops.bayesrul_b200.elbo_step(x, y, mu, log_sigma.exp(), net, mode, guide, particles, prior_loc,
                                                              prior_scale, dataset_size, seed, backend)
        ctx.save_for_backward(g_mu, g_ls)
        return sc[0].to(mu.dtype)

    @staticmethod
    def backward(ctx, g):
        g_mu, g_ls = ctx.saved_tensors
        return (g * g_mu, g * g_ls) + (None,) * 11


def elbo_loss(mu, log_sigma, x, y, *, net="inception", mode="lrt", guide="normal", particles=1, prior_loc=0.0, prior_scale=1.0,
              dataset_size: int, seed: int = 0, backend: str = "simt") -> torch.Tensor:
    """Differentiable `svi.step` loss (bayesian.py:111-132): `elbo_loss(mu, log_sigma, x, y, ...).backward()` leaves
    d loss / d mu and d loss / d log_sigma (Pyro's unconstrained scale parameter) in `.grad`, ready for any torch optimiser."""
    return _ElboLoss.apply(mu, log_sigma, x, y, net, mode, guide, particles, prior_loc, prior_scale, dataset_size, seed, backend)
