"""TEST INFRASTRUCTURE ONLY — CPU restatement (torch, any dtype) of the "msg1 by linearity" formulation that
scalable-e3-gnn_b200/csrc/msg_table.cu implements, written with the SAME index arithmetic as the kernels
(expand -> node GEMM -> per-edge combine + gate; backward: gate VJP -> transposed combine -> segment sums ->
node GEMMs -> contract).  tests/test_oracle_msg_table.py pins it to the reference-pinned ``L1TPPort``
(L1TP:242-297 on cat(x[dst], x[src], extra)) + gate with autograd; the GPU tests then compare the kernels with the
model oracle.  Only tests/ may import it.
"""
from __future__ import annotations

import math

import torch

C3 = 1.0 / math.sqrt(3.0)


def dims(ns, nv):
    return dict(MZ=ns + nv, CH=ns + 2 * nv, DPRE=ns + 4 * nv, DPOST=ns + 3 * nv, D=ns + 3 * nv, HALF=4 * (ns + 2 * nv),
                LDT=8 * (ns + 2 * nv))


def expand(ns, nv, wz, wv, nz, nvn):
    """wbig [D, LDT], we [2, CH] — msg1_expand_kernel."""
    d = dims(ns, nv)
    MZ, CH, HALF, LDT, D = d["MZ"], d["CH"], d["HALF"], d["LDT"], d["D"]
    NSC = 2 * ns + 2
    wbig = torch.zeros((D, LDT), dtype=wz.dtype)
    we = torch.zeros((2, CH), dtype=wz.dtype)
    for j in range(2):
        row = 2 * ns + j
        we[j, :MZ] = nz[:MZ] * wz[row]
        we[j, MZ:] = nvn[0::3] * C3 * wv[row]
    for k in range(D):
        for p in range(2):
            if k < ns:
                row, j, f = p * ns + k, 0, 1.0
            else:
                kv, comp = divmod(k - ns, 3)
                row, j, f = NSC + p * nv + kv, 1 + comp, C3
            base = p * HALF
            wbig[k, base + 4 * torch.arange(MZ) + j] = f * nz[:MZ] * wz[row]
            wbig[k, base + 4 * (MZ + torch.arange(nv)) + j] = C3 * nvn[0::3] * wv[row]
    return wbig, we


def edge_forward(ns, nv, table, we, y, extra, dst, src, cs, cg):
    d = dims(ns, nv)
    MZ, CH, HALF = d["MZ"], d["CH"], d["HALF"]
    E = len(dst)
    td = table[dst, :HALF].reshape(E, CH, 4)
    ts = table[src, HALF:].reshape(E, CH, 4)
    t = td + ts
    P = t[:, :, 0] + extra[:, 0:1] * we[0] + extra[:, 1:2] * we[1]
    U = t[:, :, 1:4]
    z = y[:, 0:1] * P[:, :MZ] + (y[:, None, 1:4] * U[:, :MZ]).sum(-1)
    v = y[:, None, 1:4] * P[:, MZ:, None] + y[:, 0:1, None] * U[:, MZ:]
    pre = torch.cat([z, v.reshape(E, 3 * nv)], 1)
    sg = torch.sigmoid(z)
    post = torch.cat([cs * z[:, :ns] * sg[:, :ns], (cg * sg[:, ns:, None] * v).reshape(E, 3 * nv)], 1)
    return pre, post


def edge_backward(ns, nv, pre, gpost, y, extra, dst, src, n_all, cs, cg):
    """-> gpre [E, DPRE], G [n_all, LDT], gwe [2, CH] — msg1_edge_bwd_kernel<false/true> (+ partial sums)."""
    d = dims(ns, nv)
    MZ, CH, HALF, LDT = d["MZ"], d["CH"], d["HALF"], d["LDT"]
    E = len(dst)
    z, v = pre[:, :MZ], pre[:, MZ:].reshape(E, nv, 3)
    sg = torch.sigmoid(z)
    gs, gv = gpost[:, :ns], gpost[:, ns:].reshape(E, nv, 3)
    gz_s = cs * gs * sg[:, :ns] * (1 + z[:, :ns] * (1 - sg[:, :ns]))
    dot = (gv * v).sum(-1)
    gz_g = cg * sg[:, ns:] * (1 - sg[:, ns:]) * dot
    q = cg * sg[:, ns:, None] * gv
    gz = torch.cat([gz_s, gz_g], 1)
    gpre = torch.cat([gz, q.reshape(E, 3 * nv)], 1)
    g4 = torch.zeros((E, CH, 4), dtype=pre.dtype)
    g4[:, :MZ, :] = y[:, None, :] * gz[:, :, None]
    g4[:, MZ:, 0] = (y[:, None, 1:4] * q).sum(-1)
    g4[:, MZ:, 1:4] = y[:, 0:1, None] * q
    G = torch.zeros((n_all, LDT), dtype=pre.dtype)
    G[:, :HALF].index_add_(0, dst.long(), g4.reshape(E, HALF))
    G[:, HALF:].index_add_(0, src.long(), g4.reshape(E, HALF))
    gwe = torch.stack([(extra[:, 0:1] * g4[:, :, 0]).sum(0), (extra[:, 1:2] * g4[:, :, 0]).sum(0)])
    return gpre, G, gwe


def contract(ns, nv, gwbig, gwe, nz, nvn):
    """gwz [(2ns+2+2nv), MZ], gwv [(same), nv] — msg1_contract_kernel."""
    d = dims(ns, nv)
    MZ, CH, HALF, LDT = d["MZ"], d["CH"], d["HALF"], d["LDT"]
    NSC = 2 * ns + 2
    rows = NSC + 2 * nv
    g = torch.zeros((rows, CH), dtype=gwbig.dtype)
    ch4 = 4 * torch.arange(CH)
    for row in range(rows):
        if row < 2 * ns:
            p, k = divmod(row, ns)
            g[row] = gwbig[k, p * HALF + ch4]
        elif row < NSC:
            g[row] = gwe[row - 2 * ns]
        else:
            p, kv = divmod(row - NSC, nv)
            acc = sum(gwbig[ns + 3 * kv + c, p * HALF + ch4 + 1 + c] for c in range(3))
            acc = acc.clone()
            acc[:MZ] *= C3
            g[row] = acc
    return g[:, :MZ] * nz[:MZ], g[:, MZ:] * C3 * nvn[0::3]


def msg1_forward_backward(ns, nv, x, wz, wv, nz, nvn, y, extra, dst, src, gpost, cs, cg):
    """Whole chain as the product runs it; returns (post, gx, gwz, gwv)."""
    wbig, we = expand(ns, nv, wz, wv, nz, nvn)
    table = x @ wbig
    pre, post = edge_forward(ns, nv, table, we, y, extra, dst, src, cs, cg)
    gpre, G, gwe = edge_backward(ns, nv, pre, gpost, y, extra, dst, src, x.shape[0], cs, cg)
    gx = G @ wbig.t()
    gwz, gwv = contract(ns, nv, x.t() @ G, gwe, nz, nvn)
    return post, gx, gwz, gwv
