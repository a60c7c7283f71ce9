"""Multi-GPU plumbing (torch.distributed; NCCL on GPUs, gloo in the CPU tests).

Weight gradients live in one flat fp32 buffer, so a step needs exactly one collective for them: a sum in the
Morton-range domain decomposition (``se3gnn_b200.domain``, the default for N > 1) or a mean in the data-parallel mode
(one independent cloud per rank).  ``broadcast_params_`` makes the replicas start from rank 0's weights.
"""
from __future__ import annotations

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


def _active(group=None) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    if _active(group):
        dist.all_reduce(flat, group=group)
        flat.div_(dist.get_world_size(group))
    return flat


def broadcast_params_(params, group=None) -> None:
    """Every rank of ``group`` takes the parameters of the group's first rank (one flat broadcast)."""
    if not _active(group):
        return
    params = list(params)
    flat = torch.cat([p.detach().reshape(-1) for p in params])
    dist.broadcast(flat, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    o = 0
    with torch.no_grad():
        for p in params:
            p.copy_(flat[o:o + p.numel()].view_as(p))
            o += p.numel()
