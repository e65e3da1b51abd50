"""Deep-ensemble mixture moments (reference: bayesrul/models/deepens.py:21-24, the only part of that file on the hot path).

`mixture_moments` is the kernel-facing op (brl_mixture_moments: one pass over the [M, n] member tables, double
accumulation, biased mixture variance).  `deep_ensemble(df)` is a thin pandas adapter with the reference's signature so that
`results/predictions.py:42-55` can call it unchanged; the ensemble generator / combinatorics around it are host bookkeeping
that SURVEY.md section 2 marks out of scope and stay in the reference.  There is no CPU fallback: a CUDA device is required.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import pandas as pd
import torch

from .. import _lib


def mixture_moments(mu_m, sigma_m, device=None):
    """[M, n] member means / standard deviations (array-likes or tensors) -> (mu[n], sigma[n]) as numpy arrays:
    mu = mean_m mu_m, sigma = sqrt(mean_m(mu_m^2 + sigma_m^2) - mu^2).  brl_mixture_moments takes no context, so no Engine
    (streams, events, tcgen05 state) is built for it."""
    if not torch.cuda.is_available():
        raise RuntimeError("bayesrul_b200: no CUDA device available (B200 / sm_100a required, no CPU fallback)")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    lib = _lib.load()
    mu_t = torch.as_tensor(np.ascontiguousarray(mu_m), dtype=torch.float32).to(dev).contiguous()
    sd_t = torch.as_tensor(np.ascontiguousarray(sigma_m), dtype=torch.float32).to(dev).contiguous()
    if mu_t.dim() != 2 or mu_t.shape != sd_t.shape:
        raise RuntimeError("bayesrul_b200: mu_m / sigma_m must both be [M,n]")
    M, n = mu_t.shape
    mu, sd = torch.empty(n, device=dev), torch.empty(n, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.brl_mixture_moments(mu_t.data_ptr(), sd_t.data_ptr(), M, n, mu.data_ptr(), sd.data_ptr(),
                                           C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return mu.cpu().numpy(), sd.cpu().numpy()


def deep_ensemble(df: pd.DataFrame, device=None) -> pd.DataFrame:
    """Frame with columns model / preds / labels / stds (the base learners' test outputs, one block of rows per model, same
    window order in every block) -> one frame preds / labels / stds of the mixture."""
    order = df.groupby("model").cumcount()
    wide = df.assign(_row=order).pivot(index="_row", columns="model", values=["preds", "stds"])
    mu, sigma = mixture_moments(wide["preds"].to_numpy().T, wide["stds"].to_numpy().T, device)
    labels = df.loc[df["model"] == wide["preds"].columns[0], "labels"].to_numpy()
    return pd.DataFrame({"preds": mu, "labels": labels, "stds": sigma})
