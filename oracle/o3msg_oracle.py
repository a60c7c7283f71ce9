"""TEST INFRASTRUCTURE ONLY (tests/, smoke(), the CPU arms of bench.py) - never imported by the product path.

CPU restatement (torch fp64) of the l <= 2 message product by linearity that csrc/o3msg.cu + se3gnn_b200/o3msg.py
implement:  TP(cat(x[dst], x[src], extra), Y)  ==  per-node tables  T[p][n] = x_h[n] W_p  (weight contraction without Y,
once per node and role) followed by the per-edge coupling with Y.  Pinned to oracle/lmax2_oracle.forward on the
concatenated row (tests/test_oracle_o3msg.py), which restates the reference's L1TensorProduct.forward
(/root/reference/models/segnn/l1_tensor_prod.py:242-297) for l <= 2 (for l <= 1 pinned to the reference's golden vectors).
"""
from __future__ import annotations

import torch

from . import lmax2_oracle as O2


def node_tables(x, weights, hidden, extras, in2, out):
    """x [N, dim(hidden)] -> {path index: T [N, mul_out, 2 l1 + 1]} for the paths of both roles (dst: in1 irreps
    0 .. nh-1, src: nh .. 2nh-1); no second input in it."""
    in1 = list(hidden) + list(hidden) + list(extras)
    nh = len(hidden)
    off = O2._offsets(hidden)
    tabs = {}
    for k, ((i1, i2, io), W) in enumerate(zip(O2.paths(in1, in2, out), weights)):
        if i1 >= 2 * nh:
            continue
        h = i1 % nh
        mul, l, _ = hidden[h]
        xh = x[:, off[h]:off[h] + mul * (2 * l + 1)].reshape(-1, mul, 2 * l + 1)
        tabs[k] = torch.einsum("nui,uw->nwi", xh, W)
    return tabs


def edge_couple(tabs, dst, src, extra, y, weights, hidden, extras, in2, out):
    """pre [E, dim(out)] from the node tables: gather the two roles' rows, couple with Y; the scalar extras' paths are
    contracted per edge (they have no node)."""
    in1 = list(hidden) + list(hidden) + list(extras)
    nh = len(hidden)
    a = O2.norm_factors(in1, in2, out)
    o2, ox = O2._offsets(in2), O2._offsets(extras)
    E = y.shape[0]
    res = [torch.zeros((E, mo, 2 * lo + 1), dtype=y.dtype) for mo, lo, _ in out]
    for k, ((i1, i2, io), W) in enumerate(zip(O2.paths(in1, in2, out), weights)):
        l2 = in2[i2][1]
        l1 = in1[i1][1]
        C = torch.from_numpy(O2.cg(l1, l2, out[io][1])).to(y.dtype)
        yy = y[:, o2[i2]:o2[i2] + 2 * l2 + 1]
        if i1 < 2 * nh:
            t = tabs[k][dst if i1 < nh else src]                       # [E, mul_out, d1]
        else:
            xi = i1 - 2 * nh
            t = (extra[:, ox[xi]:ox[xi] + extras[xi][0]] @ W)[:, :, None]   # l1 = 0
        res[io] = res[io] + a[io] * torch.einsum("ewi,ej,ijk->ewk", t, yy, C)
    return torch.cat([r.reshape(E, -1) for r in res], 1)


def message(x, dst, src, extra, y, weights, hidden, extras, in2, out):
    return edge_couple(node_tables(x, weights, hidden, extras, in2, out), dst, src, extra, y, weights, hidden, extras, in2, out)
