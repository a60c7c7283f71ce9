"""TEST INFRASTRUCTURE ONLY — the whole hot path on the CPU, as bench.py's reference arm times it:
numba octree graph build (self-authored spec) + SEGNN forward/backward through the torch port of the
reference's L1TensorProduct.  Label for every number produced with it: "reference TP op sequence
(port) + self-authored remainder", because only the TP exists in the reference mount."""
from __future__ import annotations

import time

import numpy as np
import torch

from . import octree_oracle as T
from .segnn_oracle import SEGNNOracle, graph_features


class CpuStep:
    def __init__(self, num_layers=4, hidden="34x0e+10x1o", threads=None, seed=0, leaf_size=32):
        if threads:
            torch.set_num_threads(int(threads))
        torch.manual_seed(seed)
        self.model = SEGNNOracle(hidden=hidden, num_layers=num_layers)
        self.opt = torch.optim.Adam(self.model.parameters(), lr=1e-3)
        self.leaf_size = leaf_size

    def step(self, pos, vel, mass, target):
        n = len(pos)
        g = T.build_graph(pos, leaf_size=self.leaf_size)
        f = graph_features(g, pos, vel, mass)
        t32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        out = self.model(t32(f["x_in"]), t32(f["node_attr"]), t32(f["edge_attr"]), t32(f["edge_extra"]),
                         torch.from_numpy(g["dst"]), torch.from_numpy(g["col"]))
        tgt = t32(target[g["order"]])
        loss = (out[:n] - tgt).square().mean()
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        return float(loss.detach()), len(g["col"])


def time_cpu(n, kind="plummer", seed=1, steps=1, warmup=1, threads=None, make_cloud=None):
    """particles/s of the CPU path on a cloud of n particles; returns (particles_per_s, ms_per_step, edges)."""
    pos, vel, mass, target = make_cloud(n, kind, seed)
    cs = CpuStep(threads=threads)
    for _ in range(warmup):
        cs.step(pos, vel, mass, target)
    t0 = time.perf_counter()
    e = 0
    for _ in range(steps):
        _, e = cs.step(pos, vel, mass, target)
    dt = (time.perf_counter() - t0) / steps
    return n / dt, dt * 1e3, e
