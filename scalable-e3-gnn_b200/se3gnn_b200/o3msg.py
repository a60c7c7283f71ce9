"""First tensor product of the l <= 2 SEGNN message layer by linearity (csrc/o3msg.cu; the l <= 1 version is ``msg.py``).

``TP(cat(x[dst], x[src], extra), Y)`` is linear in its first input, and each path is "contract the raw channels with the
weights" followed by "couple with Y".  The contraction therefore runs once per NODE — with the l <= 2 tensor-product
operator itself on a scalar second input: ``T_role = O3TP(x, 1; W_role)`` where ``W_role`` are the message product's own
weight blocks of that role, regrouped by input irrep — on ~16x fewer rows than there are edges; per edge what is left
is a gather of two table rows and the coupling with SH(2).  The backward is the transpose: coupling^T of the cotangent,
summed per destination (CSR rows) and per source (transposed order) without atomics, then the operator's own backward
at node level, which autograd routes into the SAME flat weight the tensor product owns (``state_dict`` unchanged).

Op chain replaced: ``O3TensorProduct.forward_cat([(x, dst), (x, src), (extra, None)], Y)``, i.e. ``L1TensorProduct.forward``
(/root/reference/models/segnn/l1_tensor_prod.py:242-297) generalised to l = 2 on the concatenated row.
"""
from __future__ import annotations

import math

import torch

from . import capi
from .irreps import Irreps
from .msg import EdgeIndex
from .o3tp import O3TensorProduct, _O3tpFn

__all__ = ["O3MessageTables", "supported", "tables_for"]


def _offsets(irreps):
    out, acc = [], 0
    for mi in irreps:
        out.append(acc)
        acc += mi.mul * (2 * mi.ir.l + 1)
    return out, acc


class _EdgeFn(torch.autograd.Function):
    """(tdst [n_dst, ldt], tsrc [n_all, ldt], weight, y [E, d2], extra [E, dx]) -> pre-activation [E, d_out]"""

    @staticmethod
    def forward(ctx, tdst, tsrc, weight, y, extra, lin, ei: EdgeIndex):
        dev = y.device
        pre = torch.empty((ei.e, lin.d_out), device=dev, dtype=torch.float32)
        with capi.mark("o3msg.edge_fwd", 4.0 * (ei.e * (lin.d_out + y.shape[1] + extra.shape[1] + 2 + lin.ldt) + ei.n_dst * lin.ldt)):
            capi.check(capi.lib().se3_o3msg_edge_forward(lin.io_arr, lin.nio, ei.e, capi.ptr(ei.dst), capi.ptr(ei.src),
                                                         capi.ptr(tdst), capi.ptr(tsrc), lin.ldt, capi.ptr(y), y.stride(0),
                                                         capi.ptr(extra), extra.stride(0), capi.ptr(weight), capi.ptr(pre),
                                                         lin.d_out, capi.current_stream_ptr()), "se3_o3msg_edge_forward")
        ctx.save_for_backward(y, extra)
        ctx.lin, ctx.ei, ctx.nw = lin, ei, weight.numel()
        return pre

    @staticmethod
    def backward(ctx, gpre):
        y, extra = ctx.saved_tensors
        lin, ei = ctx.lin, ctx.ei
        dev = y.device
        gpre = gpre.contiguous()
        gdst = torch.empty((ei.n_dst, lin.ldt), device=dev, dtype=torch.float32)
        gsrc = torch.empty((ei.n_all, lin.ldt), device=dev, dtype=torch.float32)
        gex = torch.empty((ei.n_dst, lin.ldg), device=dev, dtype=torch.float32) if lin.ldg else None
        with capi.mark("o3msg.edge_bwd", 4.0 * (ei.e * (2 * lin.d_out + 2 * y.shape[1] + extra.shape[1] + 1) +
                                              (ei.n_dst + ei.n_all) * lin.ldt + ei.n_dst * lin.ldg)):
            capi.check(capi.lib().se3_o3msg_edge_backward(lin.io_arr, lin.nio, ei.n_dst, ei.n_all, capi.ptr(ei.rowptr),
                                                          capi.ptr(ei.tptr), capi.ptr(ei.perm), capi.ptr(y), y.stride(0),
                                                          capi.ptr(extra), extra.stride(0), capi.ptr(gpre), lin.d_out,
                                                          capi.ptr(gdst), capi.ptr(gsrc), lin.ldt, capi.ptr(gex), max(lin.ldg, 1),
                                                          capi.current_stream_ptr()), "se3_o3msg_edge_backward")
        gw = torch.zeros(ctx.nw, device=dev, dtype=torch.float32)
        if gex is not None:
            gw[lin.gx_index(dev)] = gex.sum(0)
        return gdst, gsrc, gw, None, None, None, None


def _layout(tp: O3TensorProduct, hidden: Irreps, extra: Irreps):
    """Path bookkeeping shared by ``supported`` and the constructor; returns None if the layout is not covered."""
    H, X = list(Irreps(hidden)), list(Irreps(extra))
    nh = len(H)
    in1 = list(tp.iri1)
    if [(m.mul, m.ir.l, m.ir.p) for m in in1] != [(m.mul, m.ir.l, m.ir.p) for m in H + H + X]:
        return None
    if len({(m.ir.l, m.ir.p) for m in H}) != nh or any(m.ir.l != 0 or m.mul > 4 for m in X) or len(tp.iro) > 8:
        return None
    paths = tp._plan.paths                       # (i1, i2, io, woff, a), enumeration order (io, i2, i1)
    role = [[[(i2, io, woff) for (i1, i2, io, woff, _) in paths if i1 == r * nh + h] for h in range(nh)] for r in range(2)]
    if any([(a, b) for a, b, _ in role[0][h]] != [(a, b) for a, b, _ in role[1][h]] for h in range(nh)):
        return None
    xpaths = [(i1 - 2 * nh, i2, io, woff) for (i1, i2, io, woff, _) in paths if i1 >= 2 * nh]
    for io in range(len(tp.iro)):
        if sum(1 for h in range(nh) for (_, o, _) in role[0][h] if o == io) > capi.O3MSG_MAXP:
            return None
        xs = [p for p in xpaths if p[2] == io]
        if len(xs) > capi.O3MSG_MAXX or sum(X[p[0]].mul for p in xs) > 4 or tp.iro[io].mul > 256:
            return None
    return H, X, nh, paths, role, xpaths


def supported(tp: O3TensorProduct, hidden, extra) -> bool:
    return _layout(tp, Irreps(hidden), Irreps(extra)) is not None


class O3MessageTables:
    """Evaluates ``tp.forward_cat([(x, dst), (x, src), (extra, None)], y)`` through node tables.  Holds no parameters and no
    reference to ``tp``: the caller passes ``tp.weight``; instances are shared per irreps layout (``tables_for``)."""

    def __init__(self, tp: O3TensorProduct, hidden, extra):
        lay = _layout(tp, Irreps(hidden), Irreps(extra))
        if lay is None:
            raise capi.Se3Error("O3MessageTables: in1 must be hidden + hidden + scalar extras with distinct hidden irrep types")
        H, X, nh, paths, role, xpaths = lay
        off2, _ = _offsets(tp.iri2)
        offo, self.d_out = _offsets(tp.iro)
        offx, _ = _offsets(X)
        a_io = {io: a for (_, _, io, _, a) in paths}
        # table row: per hidden irrep h a block of N_h columns x (2 l + 1) components
        col0, nh_cols, base, acc = {}, [], [], 0
        for h in range(nh):
            c = 0
            for (i2, io, _) in role[0][h]:
                col0[(h, i2, io)] = c
                c += tp.iro[io].mul
            nh_cols.append(c)
            base.append(acc)
            acc += c * (2 * H[h].ir.l + 1)
        self.ldt = acc
        node_out = Irreps([(nh_cols[h], str(H[h].ir)) for h in range(nh) if nh_cols[h] > 0])
        self.node_tp = O3TensorProduct(Irreps(H), node_out, Irreps("1x0e"))
        # node weights = the message product's own blocks regrouped: W_node[h][u, col0 + w] = sqrt(mul_h) W_p[u, w]
        # (the operator divides by sqrt(fan-in) = sqrt(mul_h) and couples with (l, 0, l) = identity / sqrt(2l+1) * sqrt(2l+1))
        hs = [h for h in range(nh) if nh_cols[h] > 0]
        self._idx, self._scale = [], []
        for r in range(2):
            idx = torch.zeros(self.node_tp.weight.numel(), dtype=torch.int64)
            sc = torch.zeros(self.node_tp.weight.numel(), dtype=torch.float32)
            for k, ins in enumerate(self.node_tp.instructions):
                h = hs[ins.i_out]
                assert ins.i_in1 == h and ins.path_shape == (H[h].mul, 1, nh_cols[h])
                w0 = self.node_tp.weight_offsets[k]
                for (i2, io, woff) in role[r][h]:
                    mo = tp.iro[io].mul
                    u = torch.arange(H[h].mul).view(-1, 1)
                    w = torch.arange(mo).view(1, -1)
                    idx[(w0 + u * nh_cols[h] + col0[(h, i2, io)] + w).reshape(-1)] = (woff + u * mo + w).reshape(-1)
                    sc[(w0 + u * nh_cols[h] + col0[(h, i2, io)] + w).reshape(-1)] = math.sqrt(H[h].mul)
            self._idx.append(idx)
            self._scale.append(sc)
        # per output irrep descriptors
        ios = []
        gx_pairs, gacc = [], 0
        for io, mo in enumerate(tp.iro):
            ps = [(h, i2) for h in range(nh) for (i2, o, _) in role[0][h] if o == io]
            xs = [p for p in xpaths if p[2] == io]
            if not ps and not xs:
                continue
            d = capi.O3MsgIO()
            d.l, d.mul, d.off, d.a, d.np, d.nx = mo.ir.l, mo.mul, offo[io], a_io[io], len(ps), len(xs)
            for k, (h, i2) in enumerate(ps):
                d.p_l1[k], d.p_l2[k], d.p_yoff[k] = H[h].ir.l, tp.iri2[i2].ir.l, off2[i2]
                d.p_tbase[k] = base[h] + col0[(h, i2, io)] * (2 * H[h].ir.l + 1)
            slots = 0
            for k, (xi, i2, _, woff) in enumerate(xs):
                d.x_l2[k], d.x_yoff[k], d.x_woff[k], d.x_off[k], d.x_mul[k] = tp.iri2[i2].ir.l, off2[i2], woff, offx[xi], X[xi].mul
                for u in range(X[xi].mul):
                    for w in range(mo.mul):
                        gx_pairs.append((gacc + w * sum(X[p[0]].mul for p in xs) + slots + u, woff + u * mo.mul + w))
                slots += X[xi].mul
            d.gx_off, d.gx_slots = gacc, slots
            gacc += mo.mul * slots
            ios.append(d)
        self.nio = len(ios)
        self.io_arr = (capi.O3MsgIO * max(1, self.nio))(*ios)
        self.ldg = gacc
        gxi = torch.zeros(gacc, dtype=torch.int64)
        for pos, wi in gx_pairs:
            gxi[pos] = wi
        self._gx = gxi
        self._dev_cache = {}

    def _on(self, dev):
        c = self._dev_cache.get(dev)
        if c is None:
            c = self._dev_cache[dev] = ([t.to(dev) for t in self._idx], [t.to(dev) for t in self._scale], self._gx.to(dev))
        return c

    def gx_index(self, dev):
        return self._on(dev)[2]

    def __call__(self, w: torch.Tensor, xe: torch.Tensor, y: torch.Tensor, extra: torch.Tensor, ei: EdgeIndex) -> torch.Tensor:
        """w: the message product's flat weight; xe [n_all, d_hidden] (owned rows first, then halo), y [E, d_in2],
        extra [E, d_extra] -> pre-activation [E, d_out]"""
        if not (xe.is_cuda and y.is_cuda and extra.is_cuda):
            raise capi.Se3Error("O3MessageTables runs on CUDA tensors only (no CPU fallback)")
        dev = xe.device
        idx, scale, _ = self._on(dev)
        xe = xe.contiguous()
        ones = torch.ones((ei.n_all, 1), device=dev, dtype=torch.float32)
        tdst = _O3tpFn.apply(xe[:ei.n_dst], ones[:ei.n_dst], w[idx[0]] * scale[0], self.node_tp)
        tsrc = _O3tpFn.apply(xe, ones, w[idx[1]] * scale[1], self.node_tp)
        return _EdgeFn.apply(tdst, tsrc, w, y.contiguous(), extra.contiguous(), self, ei)


_TABLES: dict = {}


def tables_for(tp: O3TensorProduct, hidden, extra) -> O3MessageTables:
    key = (str(tp.iri1), str(tp.iri2), str(tp.iro), str(Irreps(hidden)), str(Irreps(extra)))
    t = _TABLES.get(key)
    if t is None:
        t = _TABLES[key] = O3MessageTables(tp, hidden, extra)
    return t
