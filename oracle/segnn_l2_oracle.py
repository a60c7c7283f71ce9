"""TEST INFRASTRUCTURE ONLY — fp64 CPU specification of the SEGNN l_max = 2 model (`models/segnn/segnn_l2.py`): the layer
layout of `oracle/segnn_oracle.py` with every tensor product evaluated by `oracle/lmax2_oracle.forward`.  Parity
unpinned (no reference source for the model; no l = 2 code in the reference at all)."""
from __future__ import annotations

import torch

from . import lmax2_oracle as O2
from .l1tp_oracle import parse_irreps
from .segnn_oracle import SIGMOID_CST, SILU_CST


class O3TPPort(torch.nn.Module):
    def __init__(self, in1, out, lmax2=2):
        super().__init__()
        self.in1, self.in2, self.out = parse_irreps(in1), O2.sh_irreps(lmax2), parse_irreps(out)
        self.shapes = O2.weight_shapes(self.in1, self.in2, self.out)
        self.weight = torch.nn.Parameter(torch.randn(sum(a * b for a, b in self.shapes), dtype=torch.float64))

    def forward(self, x1, x2):
        ws, o = [], 0
        for a, b in self.shapes:
            ws.append(self.weight[o:o + a * b].view(a, b))
            o += a * b
        return O2.forward(x1, x2, ws, self.in1, self.in2, self.out)


class SEGNNL2Oracle(torch.nn.Module):
    def __init__(self, hidden="23x0e+7x1o+4x2e", num_layers=4, out_irreps="1x1o", input_irreps="2x1o+2x0e"):
        super().__init__()
        hid = parse_irreps(hidden)
        cnt = lambda l: sum(m for m, ll, p in hid if ll == l)
        self.ns, self.nv, self.nt = cnt(0), cnt(1), cnt(2)
        self.num_layers = num_layers
        h = hidden
        hg = "+".join(f"{m}x{ir}" for m, ir in ((self.ns + self.nv + self.nt, "0e"), (self.nv, "1o"), (self.nt, "2e")) if m)
        ML = torch.nn.ModuleList
        self.embed = O3TPPort(input_irreps, h)
        self.msg1 = ML(O3TPPort(f"{h}+{h}+2x0e", hg) for _ in range(num_layers))
        self.msg2 = ML(O3TPPort(h, hg) for _ in range(num_layers))
        self.upd1 = ML(O3TPPort(f"{h}+{h}", hg) for _ in range(num_layers))
        self.upd2 = ML(O3TPPort(h, h) for _ in range(num_layers))
        self.pre1 = O3TPPort(h, hg)
        self.pre2 = O3TPPort(h, out_irreps)

    def gate(self, raw):
        ns, nv, nt = self.ns, self.nv, self.nt
        o = ns + nv + nt
        g = SIGMOID_CST * torch.sigmoid(raw[:, ns:o])
        v = raw[:, o:o + 3 * nv].reshape(-1, nv, 3) * g[:, :nv, None]
        t = raw[:, o + 3 * nv:].reshape(-1, nt, 5) * g[:, nv:, None]
        return torch.cat([SILU_CST * torch.nn.functional.silu(raw[:, :ns]), v.reshape(len(raw), -1), t.reshape(len(raw), -1)], 1)

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


def sh2_features(f, g):
    """SH(2) edge / node attributes from the l_max = 1 feature dict `f` of `oracle.segnn_oracle.graph_features` and the graph
    `g`: (edge_attr9, node_attr9) as numpy fp64 - what `se3_edge_geometry_l2` computes on the GPU."""
    import numpy as np
    P, V = f["node_pos"], f["node_vel"]
    d, s = g["dst"], g["col"]
    ea = O2.spherical_harmonics(P[s] - P[d], 2)
    na = np.zeros((len(P), 9))
    np.add.at(na, d, ea)
    na /= np.maximum(np.diff(g["rowptr"]), 1)[:, None]
    return ea, na + O2.spherical_harmonics(V, 2)
