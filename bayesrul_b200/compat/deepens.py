"""Deep-ensemble mixture moments and ensemble generator (reference: bayesrul/models/deepens.py:9-49).
`deep_ensemble(df)` keeps the pandas signature; the reduction runs on the GPU (brl_mixture_moments)
when a CUDA device is present and the frame is large, else in numpy (host bookkeeping of a few rows)."""
from __future__ import annotations

import random
from itertools import combinations
from typing import Iterator, List

import numpy as np
import pandas as pd
import torch


def mixture_moments(mu_m, sigma_m, device=None):
    """[M,n] member means / stds -> (mu[n], sigma[n]) with the biased mixture variance of deepens.py:24."""
    from ..engine import Engine

    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    eng = Engine("inception", dev)
    mu_t = torch.as_tensor(np.ascontiguousarray(mu_m), dtype=torch.float32).to(dev)
    sd_t = torch.as_tensor(np.ascontiguousarray(sigma_m), dtype=torch.float32).to(dev)
    mu, sd = eng.mixture_moments(mu_t, sd_t)
    return mu.cpu().numpy(), sd.cpu().numpy()


def deep_ensemble(df: pd.DataFrame, device=None) -> pd.DataFrame:
    labels, preds, stds = None, [], []
    for _, model in df.groupby("model"):
        if labels is None:
            labels = model.labels.values
        preds.append(model.preds.values)
        stds.append(model.stds.values)
    mu, sigma = mixture_moments(np.stack(preds), np.stack(stds), device)
    return pd.DataFrame({"preds": mu, "labels": labels, "stds": sigma})


def deep_ensemble_gen(df: pd.DataFrame, base_learners: List[str], n_models_per_ens: int, max_deepens: int,
                      device=None) -> Iterator[pd.DataFrame]:
    """deepens.py:33-49 -- `max_deepens` random `n_models_per_ens`-subsets of every base learner's runs (same
    `random.seed(1)` / `combinations` / `random.sample` sequence), each mixed by `deep_ensemble`."""
    random.seed(1)
    for method in base_learners:
        n = len(df.query(f"method=='{method}'").groupby("model"))
        comb = list(combinations(range(n), n_models_per_ens))
        for i, ens in enumerate(random.sample(comb, max_deepens)):
            models = [f"{method}_{model:03d}" for model in ens]
            yield deep_ensemble(df.query(f"model in {models}"), device).assign(method="DE", model=f"DE_{i:03d}")
