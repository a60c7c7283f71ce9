"""TEST INFRASTRUCTURE ONLY — CPU restatement of the SEGNN (l_max=1) model on the octree graph.

PARITY UNPINNED for everything except the tensor product: the reference mount holds no SEGNN
layer code, so the layer layout is the public SEGNN one (Brandstetter et al.
2021) as stated in the product's models/segnn/segnn.py; every tensor product here is the
reference-pinned ``L1TPPort``.  Submodule names match the product model so that its
``state_dict`` loads here unchanged.  Also provides the CPU graph features (numpy) that the
CUDA edge-geometry kernel is compared with, and the whole CPU pipeline that bench.py times as
the reference arm.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import octree_oracle as T
from .l1tp_port import L1TPPort

SH0 = 0.28209479177387814  # 'integral' normalisation, l=0
SH1 = 0.4886025119029199   # l=1


def _n2m(fn):
    x, w = np.polynomial.hermite.hermgauss(256)
    z = math.sqrt(2.0) * x
    return float((w * fn(z) ** 2).sum() / math.sqrt(math.pi)) ** -0.5


_sig = lambda z: 1.0 / (1.0 + np.exp(-z))
SILU_CST = _n2m(lambda z: z * _sig(z))
SIGMOID_CST = _n2m(_sig)


def graph_features(g, pos, vel, mass):
    """numpy fp64: node_pos/vel/mass, edge_attr [E,4], edge_extra [E,2], node_attr [Nn,4], x_in [Nn,8]."""
    n = g["n"]
    mm, com, cv = T.cell_moments(g, pos, vel, mass)
    o = g["order"]
    P = np.concatenate([pos[o].astype(np.float64), com])
    V = np.concatenate([vel[o].astype(np.float64), cv])
    M = np.concatenate([mass[o].astype(np.float64), mm])
    d, s = g["dst"].astype(np.int64), g["col"].astype(np.int64)
    rel = P[s] - P[d]
    r = np.linalg.norm(rel, axis=1)
    unit = np.where(r[:, None] > 0, rel / np.maximum(r, 1e-300)[:, None], 0.0)
    ea = np.concatenate([np.full((len(d), 1), SH0), SH1 * unit], 1)
    ex = np.stack([r, (n * M[d]) * (n * M[s])], 1)
    na = np.zeros((len(P), 4))
    np.add.at(na, d, ea)
    na /= np.maximum(np.diff(g["rowptr"]), 1)[:, None]
    vn = np.linalg.norm(V, axis=1)
    vu = np.where(vn[:, None] > 0, V / np.maximum(vn, 1e-300)[:, None], 0.0)
    na += np.concatenate([np.full((len(P), 1), SH0), SH1 * vu], 1)
    xin = np.concatenate([P - P[n], V, vn[:, None], (M * n)[:, None]], 1)
    return dict(node_pos=P, node_vel=V, node_mass=M, edge_attr=ea, edge_extra=ex, node_attr=na, x_in=xin)


class SEGNNOracle(torch.nn.Module):
    def __init__(self, hidden="34x0e+10x1o", num_layers=4, out_irreps="1x1o", input_irreps="2x1o+2x0e"):
        super().__init__()
        hid = T_parse(hidden)
        self.ns = sum(m for m, l, p in hid if l == 0)
        self.nv = sum(m for m, l, p in hid if l == 1)
        self.num_layers = num_layers
        h = hidden
        hg = f"{self.ns + self.nv}x0e+{self.nv}x1o"
        ML = torch.nn.ModuleList
        self.embed = L1TPPort(input_irreps, h)
        self.msg1 = ML(L1TPPort(f"{h}+{h}+2x0e", hg) for _ in range(num_layers))
        self.msg2 = ML(L1TPPort(h, hg) for _ in range(num_layers))
        self.upd1 = ML(L1TPPort(f"{h}+{h}", hg) for _ in range(num_layers))
        self.upd2 = ML(L1TPPort(h, h) for _ in range(num_layers))
        self.pre1 = L1TPPort(h, hg)
        self.pre2 = L1TPPort(h, out_irreps)

    def gate(self, raw):
        ns, nv = self.ns, self.nv
        s, g, v = raw[:, :ns], raw[:, ns:ns + nv], raw[:, ns + nv:].reshape(-1, nv, 3)
        return torch.cat([SILU_CST * torch.nn.functional.silu(s),
                          (SIGMOID_CST * torch.sigmoid(g)[:, :, None] * v).reshape(len(raw), -1)], 1)

    def forward(self, x_in, node_attr, edge_attr, edge_extra, dst, src, halo=None):
        """``halo`` (domain-decomposition tests): callable appending the halo rows to the owned rows of x."""
        dst, src = dst.long(), src.long()
        x = self.embed(x_in, node_attr)
        for l in range(self.num_layers):
            xe = x if halo is None else halo(x)
            m = self.gate(self.msg1[l](torch.cat([xe[dst], xe[src], edge_extra], 1), edge_attr))
            m = self.gate(self.msg2[l](m, edge_attr))
            agg = torch.zeros_like(x).index_add(0, dst, m)
            u = self.gate(self.upd1[l](torch.cat([x, agg], 1), node_attr))
            x = x + self.upd2[l](u, node_attr)
        return self.pre2(self.gate(self.pre1(x, node_attr)), node_attr)


def T_parse(s):
    from .l1tp_oracle import parse_irreps
    return parse_irreps(s)
