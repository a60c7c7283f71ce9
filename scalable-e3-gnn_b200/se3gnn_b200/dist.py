"""Multi-GPU plumbing (torch.distributed; NCCL on GPUs, gloo in the CPU tests).

Round 1: data parallel over independent clouds — every rank owns a whole cloud, weight gradients live in one flat
buffer and are averaged with a single all-reduce per step.  `morton_ranges` is the host logic of the next step
(Morton-range domain decomposition of one cloud): contiguous, equal-count slabs of the sorted key array.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def flatten_grads(params) -> torch.Tensor:
    """Make every p.grad a view into one flat fp32 buffer (so a step needs exactly one collective)."""
    params = list(params)
    flat = torch.zeros(sum(p.numel() for p in params), device=params[0].device, dtype=torch.float32)
    o = 0
    for p in params:
        p.grad = flat[o:o + p.numel()].view_as(p)
        o += p.numel()
    return flat


def allreduce_mean_(flat: torch.Tensor) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat)
        flat.div_(dist.get_world_size())
    return flat


def morton_ranges(n: int, world: int) -> List[Tuple[int, int]]:
    """Rank r owns Morton ranks [lo, hi): equal particle counts, remainder spread over the first ranks."""
    base, rem = divmod(n, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out
