"""Multi-GPU plumbing: member sharding and the one collective the path has.

Ensemble members are independent, so the path shards with NO data-path exchange:
rank g owns the contiguous members [M*g/G, M*(g+1)/G).  The only collective is an all-reduce of a
13-word statistics vector (sums + one max) at the end -- NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

STAT_FIELDS = ("members", "attempt", "steps", "newton_iters", "fun_evals", "jac_evals", "reject_err",
               "reject_newton", "n_failed", "flops_dense_model", "flops_executed_model")
MAX_FIELDS = ("max_attempt",)


def shard_range(M: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block partition, same formula as the C host layer (capi.cu)."""
    return (M * rank) // world, (M * (rank + 1)) // world


def pack_stats(**kw) -> dict:
    """Reduce per-member arrays (numpy or torch) to the per-rank statistics dict."""
    out = {}
    for k in STAT_FIELDS:
        v = kw.get(k, 0)
        out[k] = float(v.sum()) if hasattr(v, "sum") else float(v)
    att = kw.get("attempt", None)
    out["max_attempt"] = float(kw["max_attempt"]) if "max_attempt" in kw else (
        float(att.max()) if att is not None and hasattr(att, "max") and len(att) else 0.0)
    return out


def allreduce_stats(local: dict, device: torch.device | None = None) -> dict:
    """Sum (and max) the statistics over all ranks.  Works without an initialised process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return {k: (int(v) if float(v).is_integer() else v) for k, v in local.items()}
    dev = device if device is not None else torch.device("cpu")
    s = torch.tensor([local.get(k, 0.0) for k in STAT_FIELDS], dtype=torch.float64, device=dev)
    m = torch.tensor([local.get(k, 0.0) for k in MAX_FIELDS], dtype=torch.float64, device=dev)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    out = {k: s[i].item() for i, k in enumerate(STAT_FIELDS)}
    out.update({k: m[i].item() for i, k in enumerate(MAX_FIELDS)})
    return {k: (int(v) if float(v).is_integer() else v) for k, v in out.items()}
