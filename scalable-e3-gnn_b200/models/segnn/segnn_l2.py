"""SEGNN with l_max = 2 on the octree graph (BASELINE configs[2]), first version: every tensor product is the l <= 2 CUDA
operator (``se3gnn_b200.o3tp.O3TensorProduct``, csrc/o3tp.cu) and every gate one elementwise kernel (csrc/gate.cu); the
concatenated / gathered inputs are read in place (``forward_cat``); only the aggregation is still a torch op (the fused gather / gate / sorted-segment-sum epilogues of the l_max = 1 path, DESIGN 4.1-4.3,
are the next step for this model).  Same layer layout as ``models/segnn/segnn.py`` (public SEGNN):
embedding -> N x [message (2 gated TPs) -> add aggregation over dst -> update (gated TP, TP, residual)] -> 2 read-out TPs,
no bias terms.  Specification for the tests: ``oracle/segnn_l2_oracle.py``.

Inputs come from ``se3gnn_b200.octree``: ``g.x_in`` [Nn,8], ``g.edge_extra`` [E,2], ``g.dst`` / ``g.col`` and the SH(2)
attributes of ``octree.sh2_attributes(g)`` (edge_attr9 [E,9], node_attr9 [Nn,9]).
"""
from __future__ import annotations

import os

import torch
from torch import nn

from se3gnn_b200.gate import irreps_gate, irreps_gate_segment_sum
from se3gnn_b200.irreps import Irreps
from se3gnn_b200.o3tp import O3TensorProduct
from se3gnn_b200 import o3msg
from se3gnn_b200.msg import build_edge_index

INPUT_IRREPS = "2x1o+2x0e"   # (pos - centroid, vel, |vel|, mass)
EXTRA_IRREPS = "2x0e"        # (|rel|, m_i m_j)


def split_hidden(hidden: Irreps):
    """(ns, nv, nt) of hidden irreps of the form  a x0e + b x1o + c x2e  (the public-SEGNN BalancedIrreps layout)."""
    hidden = Irreps(hidden).simplify()
    ns, nv, nt = hidden.count("0e"), hidden.count("1o"), hidden.count("2e")
    if str(hidden) != "+".join(f"{m}x{ir}" for m, ir in ((ns, "0e"), (nv, "1o"), (nt, "2e")) if m):
        raise ValueError("hidden irreps must be of the form  a x0e + b x1o + c x2e")
    return ns, nv, nt


def gate_irreps(hidden: Irreps) -> Irreps:
    """TP output that feeds a gate: scalars, one gate scalar per l > 0 channel, then the l > 0 channels."""
    ns, nv, nt = split_hidden(hidden)
    return Irreps("+".join(f"{m}x{ir}" for m, ir in ((ns + nv + nt, "0e"), (nv, "1o"), (nt, "2e")) if m))


class SEGNNL2(nn.Module):
    def __init__(self, hidden: str = "23x0e+7x1o+4x2e", num_layers: int = 4, out_irreps: str = "1x1o",
                 input_irreps: str = INPUT_IRREPS):
        super().__init__()
        self.hidden = Irreps(hidden).simplify()
        self.ns, self.nv, self.nt = split_hidden(self.hidden)
        self.num_layers = num_layers
        h, hg, sh = str(self.hidden), gate_irreps(self.hidden), Irreps.spherical_harmonics(2)
        tp = lambda a, b: O3TensorProduct(Irreps(a), Irreps(b), sh)
        self.embed = tp(input_irreps, h)
        self.msg1 = nn.ModuleList(tp(f"{h}+{h}+{EXTRA_IRREPS}", hg) for _ in range(num_layers))
        self.msg2 = nn.ModuleList(tp(h, hg) for _ in range(num_layers))
        self.upd1 = nn.ModuleList(tp(f"{h}+{h}", hg) for _ in range(num_layers))
        self.upd2 = nn.ModuleList(tp(h, h) for _ in range(num_layers))
        self.pre1 = tp(h, hg)
        self.pre2 = tp(h, out_irreps)

    def gate(self, raw: torch.Tensor) -> torch.Tensor:
        """silu on the scalars, sigmoid gates on the l = 1 / l = 2 channels: one CUDA kernel each way (csrc/gate.cu)."""
        return irreps_gate(raw, self.ns, [(self.nv, 3), (self.nt, 5)])

    def forward(self, x_in, node_attr, edge_attr, edge_extra, dst, src, halo=None):
        """x_in [Nn,8], node_attr [Nn,9], edge_attr [E,9], edge_extra [E,2], dst/src [E] int32 (sorted by dst).

        Domain-decomposed run (``se3gnn_b200.domain``, as for the l_max = 1 model): the node arrays hold the rank's owned
        nodes, ``src`` may point past Nn into the halo and ``halo(x) -> [Nn + n_halo, d]`` appends the halo rows fetched
        from their owners (one all-to-all-v per layer, differentiable)."""
        if not x_in.is_cuda:
            raise RuntimeError("SEGNNL2 runs on CUDA (sm_100a) only; there is no CPU fallback")
        x = self.embed(x_in, node_attr)
        # message 1: "table" = weight contraction once per node by linearity (se3gnn_b200/o3msg.py), "tp" = the tensor
        # product on the concatenated row read in place (the A/B path)
        table = os.environ.get("SE3_L2_MSG", "table") == "table" and o3msg.supported(self.msg1[0], self.hidden, EXTRA_IRREPS)
        ei = None
        for l in range(self.num_layers):
            xe = x if halo is None else halo(x)
            if table:
                if ei is None:   # CSR rows by destination and the transposed order by source: once per graph
                    ei = build_edge_index(dst, src, x.shape[0], xe.shape[0])
                pre = o3msg.tables_for(self.msg1[l], self.hidden, EXTRA_IRREPS)(self.msg1[l].weight, xe, edge_attr, edge_extra, ei)
            else:
                # cat(x[dst], x[src], edge_extra) is read in place by the kernel; its gradient is scattered by the backward
                pre = self.msg1[l].forward_cat([(xe, dst, True), (xe, src), (edge_extra, None)], edge_attr)
            m = self.gate(pre)
            pre2 = self.msg2[l](m, edge_attr)
            if ei is not None:   # gate + aggregation over the sorted destinations in one kernel, no atomics
                agg = irreps_gate_segment_sum(pre2, self.ns, [(self.nv, 3), (self.nt, 5)], dst, ei.rowptr, x.shape[0])
            else:
                agg = torch.zeros_like(x).index_add_(0, dst, self.gate(pre2))
            u = self.gate(self.upd1[l].forward_cat([(x, None), (agg, None)], node_attr))
            x = x + self.upd2[l](u, node_attr)
        return self.pre2(self.gate(self.pre1(x, node_attr)), node_attr)

    def forward_graph(self, g, attrs=None):
        from se3gnn_b200.octree import sh2_attributes
        ea, na = sh2_attributes(g) if attrs is None else attrs
        return self.forward(g.x_in, na, ea, g.edge_extra, g.dst, g.col)
