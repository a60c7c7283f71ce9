"""The hot path end to end: particles -> octree graph -> SEGNN forward/backward (+ optimiser).

``TrainStep.step_device`` takes device tensors; ``TrainStep.step_host`` is the public, user-facing
call with HOST buffers (pinned staging, H2D copies and the D2H loss read inside the call).
Synthetic clouds follow SURVEY 8d (uniform cube / Plummer sphere / NFW halo, fixed seeds).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

import torch.distributed as dist

from . import domain
from .dist import allreduce_mean_, broadcast_params_, flatten_grads
from .octree import build_octree_graph


def synthetic_cloud(n: int, kind: str = "plummer", seed: int = 1):
    """(pos, vel, mass, target) float32 numpy.  target = analytic Plummer acceleration at pos."""
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        pos = rng.random((n, 3))
    else:
        d = rng.standard_normal((n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        u = rng.random(n)
        if kind == "plummer":      # a = 1, truncated at 10 a
            r = np.minimum(1.0 / np.sqrt(np.maximum(u, 1e-12) ** (-2.0 / 3.0) - 1.0), 10.0)
        elif kind == "nfw":        # c = 10, truncated at r_vir = 1: invert m(x)=ln(1+x)-x/(1+x) by bisection
            c = 10.0
            mfun = lambda x: np.log1p(x) - x / (1.0 + x)
            lo, hi = np.zeros(n), np.full(n, c)
            t = u * mfun(c)
            for _ in range(60):
                mid = 0.5 * (lo + hi)
                big = mfun(mid) > t
                hi = np.where(big, mid, hi)
                lo = np.where(big, lo, mid)
            r = 0.5 * (lo + hi) / c
        else:
            raise ValueError(kind)
        pos = r[:, None] * d
    vel = rng.standard_normal((n, 3))
    mass = np.full(n, 1.0 / n)
    r2 = (pos ** 2).sum(1, keepdims=True)
    target = -pos / (1.0 + r2) ** 1.5
    f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return f(pos), f(vel), f(mass), f(target)


class TrainStep:
    """One training step of the hot path.  ``world`` > 1: gradients are summed across ranks with one
    NCCL all-reduce of a flat buffer (each rank owns an independent cloud)."""

    def __init__(self, model, leaf_size: int = 32, lr: float = 1e-3, distributed: bool = False,
                 decompose: bool = False, group=None):
        """``distributed``: data parallel, every rank owns an independent cloud (mean of the weight gradients).
        ``decompose``: Morton-range domain decomposition of ONE cloud over the ranks (``se3gnn_b200.domain``): every
        rank is handed the whole cloud, builds the same octree, keeps the rows of the nodes it owns, exchanges halo
        rows once per layer and the weight gradients are summed."""
        self.model = model
        self.leaf_size = leaf_size
        self.distributed = distributed
        self.decompose = decompose
        self.group = group
        self.last_local = None
        params = [p for p in model.parameters()]
        if distributed or decompose:
            broadcast_params_(params, group)   # replicas must start identical: the gradients are summed / averaged
        self.flat_grad = flatten_grads(params)
        self.opt = torch.optim.Adam(params, lr=lr, fused=True)
        self._pin = None
        self.last_graph = None

    def step_device(self, pos, vel, mass, target) -> torch.Tensor:
        n = pos.shape[0]
        self.flat_grad.zero_()
        g = build_octree_graph(pos, vel, mass, leaf_size=self.leaf_size)
        self.last_graph = g
        out = self.model.forward_graph(g)
        # targets are per particle in the caller's order; nodes are in Morton-rank order
        tgt = target.index_select(0, g.order.long())
        loss = (out[:n] - tgt).square().mean()
        loss.backward()
        if self.distributed:
            allreduce_mean_(self.flat_grad, self.group)
        self.opt.step()
        return loss.detach()

    # ------------------------------------------------------------------ domain-decomposed step
    def step_device_dd(self, pos, vel, mass, target) -> torch.Tensor:
        """pos/vel/mass/target: the WHOLE cloud (identical on every rank).  Returns the global loss."""
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        n = pos.shape[0]
        self.flat_grad.zero_()
        g = build_octree_graph(pos, vel, mass, leaf_size=self.leaf_size)
        self.last_graph = g
        # ownership, local CSR, the rank's rows of the per-edge arrays and the halo lists: csrc/domain.cu, two small D2H reads
        lg = domain.local_graph_cuda(rank, world, g)
        domain.finish_halo(lg, self.group)
        self.last_local = lg
        halo = (lambda x: domain.halo_exchange(x, lg, self.group)) if world > 1 else None
        own = lg.own_ids.long()
        out = self.model(g.x_in.index_select(0, own), domain.gather_rows(g.node_attr, own), lg.edge_attr, lg.edge_extra,
                         lg.dst, lg.src, halo=halo, rowptr=lg.rowptr)
        own_part = g.order[lg.part_lo:lg.part_lo + lg.n_part].long()
        tgt = target.index_select(0, own_part)
        loss = (out[:lg.n_part] - tgt).square().sum() / (3.0 * n)
        loss.backward()
        lossd = loss.detach().clone()
        if world > 1:   # ONE collective for the weight gradients and the loss
            buf = torch.cat([self.flat_grad, lossd.reshape(1)])
            dist.all_reduce(buf, group=self.group)
            self.flat_grad.copy_(buf[:-1])
            lossd = buf[-1].clone()
        self.opt.step()
        return lossd

    def step_host_dd(self, pos_h, vel_h, mass_h, target_h) -> float:
        """Each rank passes ITS CHUNK of the cloud (pinned CPU tensors, equal sizes); the chunks are copied to the
        device and all-gathered over NCCL, then the decomposed step runs.  Returns the global loss (D2H read)."""
        dev = self.flat_grad.device
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        loc = torch.empty((pos_h.shape[0], 10), device=dev, dtype=torch.float32)
        loc[:, 0:3].copy_(pos_h, non_blocking=True)
        loc[:, 3:6].copy_(vel_h, non_blocking=True)
        loc[:, 6].copy_(mass_h, non_blocking=True)
        loc[:, 7:10].copy_(target_h, non_blocking=True)
        if world > 1:
            allp = torch.empty((world * loc.shape[0], loc.shape[1]), device=dev, dtype=loc.dtype)
            dist.all_gather_into_tensor(allp, loc, group=self.group)
        else:
            allp = loc
        pos, vel, mass, target = (allp[:, 0:3].contiguous(), allp[:, 3:6].contiguous(), allp[:, 6].contiguous(),
                                  allp[:, 7:10].contiguous())
        return float(self.step_device_dd(pos, vel, mass, target).item())

    def step_host(self, pos_h, vel_h, mass_h, target_h) -> float:
        """pos/vel/mass/target: pinned CPU tensors.  Returns the loss as a Python float (D2H read)."""
        dev = self.flat_grad.device
        pos = pos_h.to(dev, non_blocking=True)
        vel = vel_h.to(dev, non_blocking=True)
        mass = mass_h.to(dev, non_blocking=True)
        target = target_h.to(dev, non_blocking=True)
        return float(self.step_device(pos, vel, mass, target).item())
