"""Multi-GPU plumbing for the batched solve path: one process per GPU, contiguous shards, no data-path
collective (MPC instances are independent — SURVEY.md §8e).  torch.distributed (NCCL on GPUs, gloo in the
CPU tests) is used only AFTER the solve to reduce a handful of scalars: sum of {instances, solved, ADMM
iterations}, max of {four residual maxima, elapsed milliseconds}.
"""
from __future__ import annotations

import os


def shard_bounds(B: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of ceil(B/world) instances per rank; instance-major layout makes it a pointer offset."""
    per = (B + world - 1) // world
    lo = min(B, rank * per)
    return lo, min(B, lo + per)


def env_rank_world() -> tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def reduce_stats(stats: dict, device=None, group=None) -> dict:
    """All-reduce solve statistics across ranks.

    stats: {"instances", "solved", "iters": summed; "res_max": list of 4 floats, "ms": float: maxed}.
    Works with any initialised backend (nccl -> pass device=cuda:<local_rank>; gloo -> device=None/cpu).
    """
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(stats)
    dev = device if device is not None else torch.device("cpu")
    s = torch.tensor([float(stats["instances"]), float(stats["solved"]), float(stats["iters"])], dtype=torch.float64, device=dev)
    m = torch.tensor(list(stats["res_max"]) + [float(stats["ms"])], dtype=torch.float64, device=dev)
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    s, m = s.cpu().tolist(), m.cpu().tolist()
    return dict(instances=int(s[0]), solved=int(s[1]), iters=int(s[2]), res_max=m[:4], ms=m[4])
