"""Minimal stand-ins for the pyro-ppl 1.8.1 names bayesrul touches on the hot path
(bayesrul/models/bayesian.py:41,105-132,136-139,147,155,177,255-271; conf/model/bnn.yaml:6-10).

They carry configuration to ONE fused CUDA call per `svi.step`; no effect handlers, no traces.
Installed under the module names `pyro`, `pyro.infer`, `pyro.optim`, `pyro.poutine`,
`pyro.distributions` by `bayesrul_b200.compat.install_shims()` when the real packages are absent.
"""
from __future__ import annotations

import contextlib
from typing import Dict, Optional

import torch
from torch.distributions import constraints

# ---------------------------------------------------------------------------- param store


class ParamStore:
    """pyro.get_param_store(): {name: unconstrained tensor}; scales are stored as log sigma
    (constraints.positive -> exp transform, guides/radial.py:85-94)."""

    def __init__(self):
        self._params: Dict[str, torch.Tensor] = {}
        self._constraints: Dict[str, object] = {}
        self._owner = None  # the guide whose flat buffers back the views

    def clear(self):
        self._params.clear()
        self._constraints.clear()
        self._owner = None

    def bind(self, guide):
        self.clear()
        self._owner = guide
        for name, loc, log_scale in guide.named_site_views():
            self._params[name + ".loc"] = loc
            self._constraints[name + ".loc"] = constraints.real
            self._params[name + ".scale"] = log_scale
            self._constraints[name + ".scale"] = constraints.positive

    def get_state(self):
        return {"params": dict(self._params), "constraints": dict(self._constraints)}

    def set_state(self, state):
        params = state["params"]
        if self._owner is None:
            self._params = {k: v.detach().clone() for k, v in params.items()}
            self._constraints = dict(state.get("constraints", {}))
            return
        for k, v in params.items():
            if k in self._params and self._params[k].shape == v.shape:
                self._params[k].copy_(v.to(self._params[k].device))
            else:
                self._params[k] = v
        self._constraints.update(state.get("constraints", {}))

    def keys(self):
        return self._params.keys()

    def __getitem__(self, k):
        t = self._params[k]
        return t.exp() if self._constraints.get(k) is constraints.positive else t

    def pending(self):
        """state loaded before the guide existed (on_load_checkpoint precedes define_bnn)."""
        return self._params if self._owner is None else {}


_STORE = ParamStore()


def get_param_store() -> ParamStore:
    return _STORE


def clear_param_store() -> None:
    _STORE.clear()


# ---------------------------------------------------------------------------- poutine


class _Scaled:
    def __init__(self, fn, scale: float):
        self.fn, self.scale = fn, float(scale)

    def __getattr__(self, k):
        return getattr(self.fn, k)

    def __call__(self, *a, **k):
        return self.fn(*a, **k)


class _Blocked(_Scaled):
    def __init__(self, fn, hide=None):
        super().__init__(fn, getattr(fn, "scale", 1.0))
        self.hide = hide or []


class poutine:  # namespace
    @staticmethod
    def scale(fn, scale: float):
        return _Scaled(fn, scale)

    @staticmethod
    def block(fn, hide=None):
        return _Blocked(fn, hide)


# ---------------------------------------------------------------------------- ELBOs / optimiser / SVI


class TraceMeanField_ELBO:
    analytic_kl = True

    def __init__(self, num_particles: int = 1):
        self.num_particles = int(num_particles)


class Trace_ELBO(TraceMeanField_ELBO):
    analytic_kl = False


class ClippedAdam:
    """pyro.optim.ClippedAdam({lr, betas, clip_norm, lrd, weight_decay}) -- fused kernel brl_clipped_adam."""

    def __init__(self, optim_args: dict):
        self.args = dict(lr=1e-3, betas=(0.9, 0.999), eps=1e-8, clip_norm=10.0, lrd=1.0, weight_decay=0.0)
        self.args.update(optim_args)
        self.state: Dict[str, dict] = {}

    # Optimiser state is keyed by a STABLE name ("loc" / "log_scale", or the position in the list handed to step_flat) and
    # remembers which tensor it belongs to: define_bnn() builds a new guide in on_fit_start / on_test_start / on_predict_start,
    # and an `id()` of a freed tensor can be handed out again -- a new parameter must start from step 0 with zero moments.
    def _slot(self, name, p):
        self._slot_restored(name, p)
        st = self.state.get(name)
        if st is None or st["param"] is not p:
            st = dict(step=0, m=torch.zeros_like(p), v=torch.zeros_like(p), param=p)
            self.state[name] = st
        return st

    def step_flat(self, engine, params, grads, names=None):
        names = names if names is not None else [f"param{i}" for i in range(len(params))]
        for name, p, g in zip(names, params, grads):
            st = self._slot(name, p)
            st["step"] += 1
            a = self.args
            engine.clipped_adam(p, g, st["m"], st["v"], st["step"], a["lr"], tuple(a["betas"]), a["eps"], a["clip_norm"],
                                a["lrd"], a["weight_decay"])

    def step_vi(self, engine, guide, g_loc, g_log_scale):
        """The step over a mean-field guide's two flat buffers in ONE launch (brl_clipped_adam_vi), scale refresh included;
        same per-parameter state as step_flat."""
        sts = [self._slot(name, p) for name, p in (("loc", guide.loc), ("log_scale", guide.log_scale))]
        for st in sts:
            st["step"] += 1
        a = self.args
        engine.clipped_adam_vi(guide.loc, guide.log_scale, guide.scale, g_loc.contiguous(), g_log_scale.contiguous(), sts[0]["m"],
                               sts[0]["v"], sts[1]["m"], sts[1]["v"], sts[0]["step"], a["lr"], tuple(a["betas"]), a["eps"],
                               a["clip_norm"], a["lrd"], a["weight_decay"])

    def get_state(self):
        """{name: {step, m, v}} -- restorable with set_state (names, not object ids)."""
        return {k: dict(step=v["step"], m=v["m"].clone(), v=v["v"].clone()) for k, v in self.state.items()}

    def set_state(self, state, params=None):
        """Restore get_state(); `params` = {name: tensor} binds the slots to live parameters (else bound on first use)."""
        self.state = {}
        for k, v in state.items():
            self.state[k] = dict(step=int(v["step"]), m=v["m"].clone(), v=v["v"].clone(), param=(params or {}).get(k))

    def _slot_restored(self, name, p):  # a restored slot without a bound parameter adopts the first tensor of matching shape
        st = self.state.get(name)
        if st is not None and st["param"] is None and st["m"].shape == p.shape:
            st["param"] = p
            st["m"], st["v"] = st["m"].to(p.device), st["v"].to(p.device)


def _unwrap(fn):
    scale, blocked = 1.0, False
    while isinstance(fn, _Scaled):
        if isinstance(fn, _Blocked):
            blocked = True
        else:
            scale *= fn.scale
        fn = fn.fn
    return fn, scale, blocked


class SVI:
    """SVI(model, guide, optim, loss): `.step(x, y) -> float` runs forward + ELBO backward + optimiser in
    CUDA; `.evaluate_loss(...)` is the no-grad twin.  `model` is `bnn.model` (possibly wrapped by
    poutine.scale / poutine.block), `guide` is `bnn.guide` (bayesian.py:111-132,136-139)."""

    def __init__(self, model, guide, optim, loss):
        self.model_h, self.model_scale, self.no_obs = _unwrap(model)
        self.guide_h, self.guide_scale, _ = _unwrap(guide)
        self.optim, self.loss = optim, loss
        self.bnn = getattr(self.model_h, "__self__", self.model_h)
        # bayesian.py:136 blocks the bnn itself: calling bnn(x) has no likelihood -> KL only
        self.no_obs = self.no_obs or not getattr(self.model_h, "_is_model", False)
        self.last: Optional[dict] = None

    def _run(self, x, y, grads: bool):
        res = self.bnn._elbo(x, y, particles=self.loss.num_particles, analytic_kl=self.loss.analytic_kl, grads=grads)
        self.last = res
        sc = res["scalars"]
        # (device scalar, host factor): the factor is applied AFTER the .item() read -- a device multiply would be one more launch
        # on the critical path of every step
        if self.no_obs:  # unscaled KL (A.6 svi_no_obs)
            return res, sc[2], float(self.guide_scale)
        return res, sc[0], float(self.model_scale / res["c"])

    def step(self, x, y=None) -> float:
        res, loss, factor = self._run(x, y, True)
        k = self.model_scale / res["c"]
        g_mu, g_ls = res["grad_mu"], res["grad_log_sigma"]
        if k != 1.0:
            g_mu, g_ls = g_mu * k, g_ls * k
        guide = self.bnn.net_guide
        if self.optim is not None:
            if hasattr(self.optim, "step_vi") and guide.loc.is_contiguous() and guide.log_scale.is_contiguous():
                self.optim.step_vi(self.bnn.engine, guide, g_mu, g_ls)  # one launch: both buffers + scale = exp(log scale)
            else:
                self.optim.step_flat(self.bnn.engine, [guide.loc, guide.log_scale], [g_mu, g_ls], names=["loc", "log_scale"])
                guide.refresh()
        return float(loss.item()) * factor  # device boundary #2 of the reference (.item() sync every step)

    def evaluate_loss(self, x, y=None) -> float:
        if y is None and self.no_obs and self.bnn._last_kl is not None:
            return float(self.bnn._last_kl.item())  # KL does not depend on the batch for analytic KL
        res, loss, factor = self._run(x, y if y is not None else torch.zeros(x.shape[0], device=x.device), False)
        return float(loss.item()) * factor
