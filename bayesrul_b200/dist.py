"""Multi-GPU plumbing (one process per GPU, torch.distributed: NCCL on GPUs, gloo in the CPU tests).

The path shards without a data-path collective:
  * predict: contiguous blocks of windows per rank, every rank runs all S samples with the same Philox
    key -> identical weight draws everywhere, results are concatenated in rank order;
  * MC-sample / ensemble-member sharding: each rank reduces its own samples to (n, mean, M2, sum sigma^2)
    per window; one all-gather + Chan's parallel-variance merge gives the global moments;
  * training: data-parallel minibatches, ONE flat all-reduce(avg) of [grad_mu | grad_log_sigma | scalars].
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, order-preserving block of `n` items for `rank` (first n % world ranks get one more)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def merge_moments(parts: Sequence[Tuple[object, torch.Tensor, torch.Tensor, torch.Tensor]]):
    """Chan merge of per-shard predictive moments.  parts = [(n_r, pred_r, ep_var_r (unbiased), al_var_r)]; n_r is a Python
    number or a tensor (0-dim or broadcastable to pred_r: nothing is read back to the host).
    Returns (pred, std, ep_var, al_var) identical to reducing all samples at once (bayesian.py:212-215)."""
    ns = [p[0].double() if isinstance(p[0], torch.Tensor) else float(p[0]) for p in parts]
    n = sum(ns)
    mean = sum(k * p[1].double() for k, p in zip(ns, parts)) / n
    m2 = sum(p[2].double().nan_to_num(0.0) * (k - 1) + k * (p[1].double() - mean) ** 2 for k, p in zip(ns, parts))
    al = sum(k * p[3].double() for k, p in zip(ns, parts)) / n
    ep = m2 / (n - 1)  # n == 1: 0 / 0 = NaN, like loc.var(0) of a single sample
    dt = parts[0][1].dtype
    return mean.to(dt), (al + ep).sqrt().to(dt), ep.to(dt), al.to(dt)


def all_gather_moments(n_local: int, pred, ep_var, al_var, group=None):
    """Moment merge across ranks that each hold a different subset of MC samples of the SAME windows: one all-gather of
    [4, B] per rank, Chan merge on the device (no host synchronisation)."""
    world = dist.get_world_size(group)
    packed = torch.stack([torch.full_like(pred, float(n_local)), pred, ep_var.nan_to_num(0.0), al_var])
    flat = torch.empty((world * packed.shape[0],) + tuple(packed.shape[1:]), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(flat, packed, group=group)  # rank-major concatenation along dim 0
    bufs = flat.view((world,) + tuple(packed.shape))
    return merge_moments([(bufs[r, 0], bufs[r, 1], bufs[r, 2], bufs[r, 3]) for r in range(world)])


def mixture_across_ranks(mu_local: torch.Tensor, sigma_local: torch.Tensor, group=None):
    """Deep-ensemble mixture (deepens.py:21-24, biased variance) when members are spread over ranks:
    mu_local / sigma_local are [M_r, n]."""
    s = torch.stack([mu_local.double().sum(0), (mu_local.double() ** 2 + sigma_local.double() ** 2).sum(0),
                     torch.full((mu_local.shape[1],), float(mu_local.shape[0]), dtype=torch.float64, device=mu_local.device)])
    dist.all_reduce(s, group=group)
    mu = s[0] / s[2]
    return mu.to(mu_local.dtype), (s[1] / s[2] - mu**2).sqrt().to(mu_local.dtype)


def allreduce_elbo_grads(res: dict, group=None) -> dict:
    """Average the ELBO step outputs of data-parallel ranks with ONE collective over the step's flat result buffer
    [grad_mu | grad_log_sigma | loss, nll, kl, mse] (Engine.elbo_step: res["flat"]); grad_sigma is not reduced (the optimiser
    steps log sigma).  Each rank computed its loss with plate scale N/B_r, so the mean over ranks is the global-batch
    gradient (the KL term is rank-replicated); NCCL's AVG folds the 1 / world in."""
    flat = res["flat"]
    P = res["grad_mu"].numel()
    flat[2 * P: 2 * P + 4].copy_(res["scalars"])  # float64 -> the fp32 tail of the reduced range
    red = flat[: 2 * P + 4]
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(red, op=dist.ReduceOp.AVG, group=group)
    else:  # gloo (CPU tests) has no AVG
        dist.all_reduce(red, group=group)
        red /= dist.get_world_size(group)
    out = dict(res)
    out["scalars"] = flat[2 * P: 2 * P + 4].double()
    return out


class FlatGradAllReduce:
    """The data-parallel gradient exchange of one ELBO step as ONE low-latency collective.

    Owns the step's reduced result buffer `[grad_mu | grad_log_sigma | loss, nll, kl, mse]` (2 P + 4 floats, 1.5 MB for Inception): pass
    `out_flat=reducer.flat` to `Engine.elbo_step` so that the kernels write their gradients straight into it.  When torch's symmetric
    memory is available the buffer is NVLink-mapped and `reduce()` runs the multimem (NVLS, in-switch reduction) all-reduce kernel --
    measured 16.8 us at 8 x B200 against 32.9 us for ncclAllReduce (tools/probe_symm.py; the exchange is latency-bound) -- else NCCL's AVG.
    The multimem kernel SUMS: `grad_scale` (1 / world) is what the optimiser must multiply the gradients with
    (`Engine.clipped_adam_vi(..., grad_scale=reducer.grad_scale)`), the scalars are scaled here."""

    def __init__(self, P: int, device, group=None, prefer_symm: bool = True, numel: int = 0):
        """numel > 0: a plain buffer of that many floats instead of the ELBO layout (e.g. the HNN step's [P] gradient: use
        `reduce_flat()`)."""
        self.P, self.group = P, group
        self.world = dist.get_world_size(group)
        n = numel if numel > 0 else 2 * P + 4
        n_pad = (n + 7) // 8 * 8
        self.mode, self.grad_scale = "nccl", 1.0
        self.flat = None
        # measured (tools/probe_symm.py, 1.5 MB): 2 GPUs NCCL 20.7 us / multimem 22.9 us; 8 GPUs NCCL 32.9 us / multimem 16.8 us --
        # the in-switch reduction pays from four ranks on
        if prefer_symm and self.world >= 4 and dist.get_backend(group) == "nccl":
            try:
                import torch.distributed._symmetric_memory as symm_mem

                self._group_name = (group or dist.group.WORLD).group_name
                buf = symm_mem.empty(n_pad, device=device, dtype=torch.float32)
                hdl = symm_mem.rendezvous(buf, self._group_name)
                if getattr(hdl, "multicast_ptr", 0):
                    buf.zero_()
                    self.flat, self.mode, self.grad_scale = buf, "multimem", 1.0 / self.world
            except Exception:  # noqa: BLE001 -- no NVLS / no symmetric memory in this build: NCCL below
                self.flat = None
        if self.flat is None:
            self.flat = torch.zeros(n_pad, device=device)

    def reduce_flat(self) -> float:
        """All-reduce self.flat in place; returns the factor that turns its content into the mean over the ranks."""
        if self.world == 1:
            return 1.0
        if self.mode == "multimem":
            torch.ops.symm_mem.multimem_all_reduce_(self.flat, "sum", self._group_name)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
        return self.grad_scale

    def reduce(self, res: dict) -> dict:
        """res = Engine.elbo_step(..., out_flat=self.flat).  Returns res with rank-averaged scalars; the gradients in
        res["grad_mu"] / res["grad_log_sigma"] (views of self.flat) hold the SUM (multimem) or the mean (nccl) over the ranks."""
        P = self.P
        self.flat[2 * P: 2 * P + 4].copy_(res["scalars"])
        out = dict(res)
        if self.world == 1:
            return out
        if self.mode == "multimem":
            torch.ops.symm_mem.multimem_all_reduce_(self.flat, "sum", self._group_name)
            out["scalars"] = self.flat[2 * P: 2 * P + 4].double() / self.world
        else:
            dist.all_reduce(self.flat[: 2 * P + 4], op=dist.ReduceOp.AVG, group=self.group)
            out["scalars"] = self.flat[2 * P: 2 * P + 4].double()
        return out


def gather_predictions(local: Sequence[torch.Tensor], group=None) -> List[torch.Tensor]:
    """Concatenate per-rank window blocks in rank order (blocks may differ by one window)."""
    world = dist.get_world_size(group)
    n_loc = torch.tensor([local[0].shape[0]], device=local[0].device)
    sizes = [torch.zeros_like(n_loc) for _ in range(world)]
    dist.all_gather(sizes, n_loc, group=group)
    mx = int(max(s.item() for s in sizes))
    out = []
    for t in local:
        pad = torch.zeros(mx, dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
        out.append(torch.cat([b[: int(s.item())] for b, s in zip(bufs, sizes)]))
    return out
