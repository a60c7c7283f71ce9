"""Host side of the SEGNN message layer's first tensor product by linearity (csrc/msg_table.cu).

``TP(cat(x[dst], x[src], extra), Y)`` is linear in its first input, so the weight contraction runs once per NODE
(``T = x . wbig``, a dense GEMM over irrep channels with ~17x fewer rows than there are edges) and the per-edge work
is the combination with SH(1) plus the gate.  The backward is the exact transpose: gate VJP and segment sums per edge
(no atomics: CSR rows by destination, stable transposed order by source), then node-level GEMMs for the input and
weight gradients.  The reference op chain this replaces: ``L1TensorProduct.forward`` (L1TP:242-297) on the
concatenated row, followed by the public-SEGNN swish gate.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import torch

from . import capi


@dataclass
class EdgeIndex:
    """CSR views of one (local) edge list that the message kernels need, built once per graph."""
    n_dst: int                 # destination (owned) nodes
    n_all: int                 # rows of the node-feature matrix the sources index (owned + halo)
    e: int
    rowptr: torch.Tensor       # [n_dst + 1] int64, edges sorted by destination
    src: torch.Tensor          # [e] int32
    dst: torch.Tensor          # [e] int32
    tptr: torch.Tensor         # [n_all + 1] int64, edges in stable order by source
    perm: torch.Tensor         # [e] int32


def build_edge_index(dst: torch.Tensor, src: torch.Tensor, n_dst: int, n_all: int,
                     rowptr: Optional[torch.Tensor] = None) -> EdgeIndex:
    lib = capi.lib()
    dev = dst.device
    e = int(dst.numel())
    st = capi.current_stream_ptr()
    if rowptr is None or rowptr.numel() != n_dst + 1:
        rowptr = torch.empty(n_dst + 1, device=dev, dtype=torch.int64)
        capi.check(lib.se3_rowptr_from_sorted(e, n_dst, capi.ptr(dst), rowptr.data_ptr(), st), "se3_rowptr_from_sorted")
    nb = C.c_size_t()
    capi.check(lib.se3_graph_transpose_work_bytes(n_all, C.byref(nb)))
    work = torch.empty(nb.value, device=dev, dtype=torch.uint8)
    tptr = torch.empty(n_all + 1, device=dev, dtype=torch.int64)
    perm = torch.empty(max(e, 1), device=dev, dtype=torch.int32)
    with capi.mark("graph.transpose", 4.0 * e * 4 + 16.0 * n_all):
        capi.check(lib.se3_graph_transpose(e, n_all, capi.ptr(src), tptr.data_ptr(), perm.data_ptr(), work.data_ptr(),
                                           nb.value, st), "se3_graph_transpose")
    return EdgeIndex(n_dst=n_dst, n_all=n_all, e=e, rowptr=rowptr, src=src, dst=dst, tptr=tptr, perm=perm)


def _node_forward(xe, wz, wv, nz, nvn, ns, nv, n_all, tag):
    """node tables T = x . W (csrc/msg_node.cu) and the extras' folded weights `we`"""
    lib = capi.lib()
    dev = xe.device
    st = capi.current_stream_ptr()
    ch, d = ns + 2 * nv, ns + 3 * nv
    we = torch.empty((2, ch), device=dev, dtype=torch.float32)
    capi.check(lib.se3_msg1_expand(ns, nv, wz.data_ptr(), wv.data_ptr(), capi.ptr(nz), capi.ptr(nvn), None, we.data_ptr(), st),
               "se3_msg1_expand")
    table = torch.empty((n_all, 8 * ch), device=dev, dtype=torch.float32)
    with capi.mark(tag, 4.0 * n_all * (d + 8 * ch), 2.0 * n_all * 2 * ch * d):
        capi.check(lib.se3_msg1_node_table(ns, nv, n_all, xe.data_ptr(), wz.data_ptr(), wv.data_ptr(), capi.ptr(nz),
                                           capi.ptr(nvn), table.data_ptr(), st), "se3_msg1_node_table")
    return table, we


def _node_backward(xe, G, wz, wv, nz, nvn, parts, nparts, ns, nv, n_all, need_gx, tag):
    """gx = G . W^T, gwz / gwv = x^T . G (+ extras' rows), straight into the parameters' layout"""
    lib = capi.lib()
    dev = xe.device
    ch, d = ns + 2 * nv, ns + 3 * nv
    mp, pf = C.c_int32(), C.c_int32()
    capi.check(lib.se3_msg1_node_parts(ns, nv, C.byref(mp), C.byref(pf)))
    scratch = torch.empty((mp.value, pf.value), device=dev, dtype=torch.float32)
    gx = torch.empty_like(xe) if need_gx else None
    gwz, gwv = torch.empty_like(wz), torch.empty_like(wv)
    with capi.mark(tag, 4.0 * n_all * (2 * d + 2 * 8 * ch), 2.0 * 2 * n_all * 2 * ch * d):
        capi.check(lib.se3_msg1_node_backward(ns, nv, n_all, xe.data_ptr(), G.data_ptr(), wz.data_ptr(), wv.data_ptr(),
                                              capi.ptr(nz), capi.ptr(nvn), parts.data_ptr(), nparts, capi.ptr(gx),
                                              gwz.data_ptr(), gwv.data_ptr(), scratch.data_ptr(), mp.value,
                                              capi.current_stream_ptr()), "se3_msg1_node_backward")
    return gx, gwz, gwv


def supported(ns: int, nv: int, n_extra: int) -> bool:
    return bool(capi.lib().se3_msg1_supported(ns, nv, n_extra))


class Msg1Fn(torch.autograd.Function):
    """(xe [n_all, d], wz, wv, nz, nv_norm, y [E,4], extra [E,2], ei, ns, nv, cs, cg) -> gated message [E, ns + 3 nv]."""

    @staticmethod
    def forward(ctx, xe, wz, wv, nz, nvn, y, extra, ei: EdgeIndex, ns: int, nv: int, cs: float, cg: float):
        lib = capi.lib()
        dev = xe.device
        st = capi.current_stream_ptr()
        ch, d = ns + 2 * nv, ns + 3 * nv
        assert xe.shape == (ei.n_all, d) and xe.is_contiguous() and xe.dtype == torch.float32
        table, we = _node_forward(xe, wz, wv, nz, nvn, ns, nv, ei.n_all, "msg.table")
        pre = torch.empty((ei.e, ns + 4 * nv), device=dev, dtype=torch.float32)
        post = torch.empty((ei.e, d), device=dev, dtype=torch.float32)
        with capi.mark("msg1.edge_fwd", 4.0 * (ei.e * (4 + 2 + 1 + ns + 4 * nv + d) + (ei.n_all + ei.n_dst) * 4 * ch)):
            capi.check(lib.se3_msg1_edge_forward(ns, nv, ei.n_dst, ei.rowptr.data_ptr(), ei.src.data_ptr(), table.data_ptr(),
                                                 we.data_ptr(), y.data_ptr(), extra.data_ptr(), cs, cg, pre.data_ptr(),
                                                 post.data_ptr(), st), "se3_msg1_edge_forward")
        ctx.ei, ctx.dims = ei, (ns, nv, cs, cg)
        ctx.save_for_backward(xe, pre, y, extra, nz, nvn, wz, wv)
        return post

    @staticmethod
    def backward(ctx, gpost):
        lib = capi.lib()
        xe, pre, y, extra, nz, nvn, wz, wv = ctx.saved_tensors
        ei: EdgeIndex = ctx.ei
        ns, nv, cs, cg = ctx.dims
        dev = xe.device
        st = capi.current_stream_ptr()
        ch, d = ns + 2 * nv, ns + 3 * nv
        gpost = gpost.contiguous()
        gpre = torch.empty_like(pre)
        G = torch.empty((ei.n_all, 8 * ch), device=dev, dtype=torch.float32)
        parts = torch.empty((int(lib.se3_msg1_max_parts()), 2, ch), device=dev, dtype=torch.float32)
        nparts = C.c_int32()
        with capi.mark("msg1.edge_bwd", 4.0 * (ei.e * (2 * 4 + 2 + 1 + d + 3 * (ns + 4 * nv)) + 2 * ei.n_all * 4 * ch)):
            capi.check(lib.se3_msg1_edge_backward(ns, nv, ei.n_dst, ei.n_all, ei.rowptr.data_ptr(), ei.tptr.data_ptr(),
                                                  ei.perm.data_ptr(), y.data_ptr(), extra.data_ptr(), pre.data_ptr(),
                                                  gpost.data_ptr(), cs, cg, gpre.data_ptr(), G.data_ptr(), parts.data_ptr(),
                                                  C.byref(nparts), st), "se3_msg1_edge_backward")
        gx, gwz, gwv = _node_backward(xe, G, wz, wv, nz, nvn, parts, nparts.value, ns, nv, ei.n_all,
                                      ctx.needs_input_grad[0], "msg.node_bwd")
        return gx, gwz, gwv, None, None, None, None, None, None, None, None, None


def msg1(xe, wz, wv, nz, nvn, y, extra, ei: EdgeIndex, ns: int, nv: int, cs: float, cg: float) -> torch.Tensor:
    for name, t in (("x", xe), ("weights_l0e", wz), ("weights_l1o", wv), ("edge_attr", y), ("edge_extra", extra)):
        if not t.is_cuda:
            raise RuntimeError(f"se3gnn_b200.msg1: {name} must be a CUDA tensor (there is no CPU fallback)")
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise TypeError(f"se3gnn_b200.msg1: {name} must be contiguous float32")
    return Msg1Fn.apply(xe, wz, wv, nz, nvn, y, extra, ei, ns, nv, cs, cg)


# ------------------------------------------------------------------------------------------------ fused message layer
DBG_TIMING = None   # tools/bench_msg.py --timing sets this to a list: receives the phase cycle counters of every forward
# weight-gradient kernel of message 2 on a side stream: OFF.  Measured (gpurun c3): no gain end to end (16.61 vs 16.44 ms
# per step) and a slower device-timed loop; the kernel's one CTA per SM (196 KB of shared memory) leaves no room for the
# kernels it was meant to overlap with.  SE3_OVERLAP=1 re-enables it for experiments.
OVERLAP_BWDW = os.environ.get("SE3_OVERLAP", "0") == "1"
_SIDE = {}


def _side_stream(dev):
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    s = _SIDE.get(key)
    if s is None:
        s = _SIDE[key] = torch.cuda.Stream(device=dev)
    return s
def fused_supported(ns: int, nv: int, n_extra: int) -> bool:
    return bool(capi.lib().se3_msg_fused_supported(ns, nv, n_extra)) and supported(ns, nv, n_extra)


class MsgLayerFn(torch.autograd.Function):
    """The whole message layer: (xe, message-1 weights, message-2 weights, SH, extras, edge index) -> agg [n_dst, d].

    Forward: node-table GEMM + ONE fused tcgen05 kernel (csrc/msg_fused_fwd.cu).  Backward: message 2 through the
    tensor-product backward kernels (input gradient written per edge, weight gradients), message 1 through the
    segment-sum kernels of csrc/msg_table.cu and the node-level GEMMs."""

    @staticmethod
    def forward(ctx, xe, wz1, wv1, nz1, nv1, wz2, wv2, nz2, nv2, y, extra, ei: EdgeIndex, ns, nv, cs, cg, plan2):
        lib = capi.lib()
        dev = xe.device
        st = capi.current_stream_ptr()
        ch, d, dpre = ns + 2 * nv, ns + 3 * nv, ns + 4 * nv
        assert xe.shape == (ei.n_all, d) and xe.is_contiguous() and xe.dtype == torch.float32
        table, we = _node_forward(xe, wz1, wv1, nz1, nv1, ns, nv, ei.n_all, "msg.table")
        # the backward kernel stages whole 64-row blocks of the pre-activations with cp.async.bulk: rows rounded up
        epad = (ei.e + 63) // 64 * 64
        pre1 = torch.empty((epad, dpre), device=dev, dtype=torch.float32)[:ei.e]
        m1 = torch.empty((epad, d), device=dev, dtype=torch.float32)[:ei.e]
        pre2 = torch.empty((epad, dpre), device=dev, dtype=torch.float32)[:ei.e]
        agg = torch.zeros((ei.n_dst, d), device=dev, dtype=torch.float32)
        # algorithmic bytes: SH + extras + both indices, the three per-edge tensors the backward reads, the node tables
        # (each row once) and the aggregate
        nbytes = 4.0 * (ei.e * (4 + 2 + 2 + 2 * dpre + d) + (ei.n_all + ei.n_dst) * 4 * ch + ei.n_dst * d)
        flops = 2.0 * ei.e * ((ns + nv) * (ns + nv) + ns * nv + 3 * nv * nv + 3 * nv * (ns + nv))
        dbg = None
        if DBG_TIMING is not None:
            dbg = torch.zeros((148, 2, 8), device=dev, dtype=torch.int64)
            DBG_TIMING.append(dbg)
        with capi.mark("msg.fused_fwd", nbytes, flops):
            capi.check(lib.se3_msg_fused_forward_dbg(ns, nv, ei.e, ei.dst.data_ptr(), ei.src.data_ptr(), table.data_ptr(),
                                                     we.data_ptr(), y.data_ptr(), extra.data_ptr(), wz2.data_ptr(),
                                                     wv2.data_ptr(), capi.ptr(nz2), capi.ptr(nv2), cs, cg, pre1.data_ptr(),
                                                     m1.data_ptr(), pre2.data_ptr(), agg.data_ptr(), capi.ptr(dbg), st),
                       "se3_msg_fused_forward")
        ctx.ei, ctx.dims, ctx.plan2 = ei, (ns, nv, cs, cg), plan2
        ctx.save_for_backward(xe, pre1, m1, pre2, y, extra, nz1, nv1, wz1, wv1, wz2, wv2, nz2, nv2)
        return agg

    @staticmethod
    def backward(ctx, gagg):
        lib = capi.lib()
        xe, pre1, m1, pre2, y, extra, nz1, nv1, wz1, wv1, wz2, wv2, nz2, nv2 = ctx.saved_tensors
        ei: EdgeIndex = ctx.ei
        ns, nv, cs, cg = ctx.dims
        dev = xe.device
        st = capi.current_stream_ptr()
        ch, d, dpre = ns + 2 * nv, ns + 3 * nv, ns + 4 * nv
        gagg = gagg.contiguous()
        # ---- input-gradient side in ONE tcgen05 kernel: gate VJP (message 2) -> W2^T -> gate VJP (message 1)
        epad = (ei.e + 63) // 64 * 64
        gpre1 = torch.empty((epad, dpre), device=dev, dtype=torch.float32)[:ei.e]
        gpre2 = torch.empty((epad, dpre), device=dev, dtype=torch.float32)[:ei.e]
        rowb = 4.0 * (4 + 1 + dpre)
        with capi.mark("msg.fused_bwd", ei.e * (rowb + 4.0 * 3 * dpre) + 4.0 * ei.n_dst * d,
                       2.0 * ei.e * ((ns + nv) * (ns + nv) + ns * nv + 3 * nv * nv + 3 * nv * (ns + nv))):
            capi.check(lib.se3_msg_fused_backward(ns, nv, ei.e, ei.dst.data_ptr(), y.data_ptr(), pre1.data_ptr(),
                                                  pre2.data_ptr(), gagg.data_ptr(), wz2.data_ptr(), wv2.data_ptr(),
                                                  capi.ptr(nz2), capi.ptr(nv2), cs, cg, gpre1.data_ptr(), gpre2.data_ptr(), st),
                       "se3_msg_fused_backward")
        # ---- weight gradient of message 2 (csrc/msg_fused_bwdw.cu: MN-major tcgen05, accumulators resident in TMEM)
        gwz2, gwv2 = torch.empty_like(wz2), torch.empty_like(wv2)
        mp, pf = C.c_int32(), C.c_int32()
        capi.check(lib.se3_msg_fused_bwdw_parts(ns, nv, C.byref(mp), C.byref(pf)))
        wparts = torch.empty((mp.value, pf.value), device=dev, dtype=torch.float32)
        # It depends only on g_pre2 and nothing below depends on it; optionally (SE3_OVERLAP=1, see OVERLAP_BWDW) on a side
        # stream, joined before this function returns.
        side = _side_stream(dev) if (capi._prof is None and OVERLAP_BWDW) else None
        main = torch.cuda.current_stream(dev)
        if side is not None:
            side.wait_stream(main)
        with capi.mark("msg.fused_bwdw", 4.0 * ei.e * (4 + d + dpre),
                       2.0 * ei.e * ((ns + nv) * (ns + nv) + ns * nv + 3 * nv * nv + 3 * nv * (ns + nv))):
            capi.check(lib.se3_msg_fused_backward_w(ns, nv, ei.e, y.data_ptr(), m1.data_ptr(), gpre2.data_ptr(),
                                                    capi.ptr(nz2), capi.ptr(nv2), gwz2.data_ptr(), gwv2.data_ptr(),
                                                    wparts.data_ptr(), mp.value, side.cuda_stream if side is not None else st),
                       "se3_msg_fused_backward_w")
        # ---- message 1: transposed SH combine + segment sums (dst rows, then the transposed order), node-level kernels
        G = torch.empty((ei.n_all, 8 * ch), device=dev, dtype=torch.float32)
        parts = torch.empty((int(lib.se3_msg1_max_parts()), 2, ch), device=dev, dtype=torch.float32)
        nparts = C.c_int32()
        with capi.mark("msg1.edge_bwd", 4.0 * (ei.e * (2 * 4 + 2 + 1 + 2 * dpre) + 2 * ei.n_all * 4 * ch)):
            capi.check(lib.se3_msg1_edge_backward(ns, nv, ei.n_dst, ei.n_all, ei.rowptr.data_ptr(), ei.tptr.data_ptr(),
                                                  ei.perm.data_ptr(), y.data_ptr(), extra.data_ptr(), None, None, cs, cg,
                                                  gpre1.data_ptr(), G.data_ptr(), parts.data_ptr(), C.byref(nparts), st),
                       "se3_msg1_edge_backward")
        gx, gwz1, gwv1 = _node_backward(xe, G, wz1, wv1, nz1, nv1, parts, nparts.value, ns, nv, ei.n_all,
                                        ctx.needs_input_grad[0], "msg.node_bwd")
        if side is not None:
            main.wait_stream(side)      # gwz2 / gwv2 (and the scratch they were built from) are complete from here on
        return (gx, gwz1, gwv1, None, None, gwz2, gwv2, None, None) + (None,) * 8


def message_layer(xe, w1, n1, w2, n2, y, extra, ei: EdgeIndex, ns: int, nv: int, cs: float, cg: float, plan2) -> torch.Tensor:
    """w1 / w2 = (weights_l0e, weights_l1o), n1 / n2 = (norm_l0e, norm_l1o) of the two message tensor products."""
    for name, t in (("x", xe), ("edge_attr", y), ("edge_extra", extra), *[(f"weight{i}", w) for i, w in enumerate(w1 + w2)]):
        if not t.is_cuda:
            raise RuntimeError(f"se3gnn_b200.message_layer: {name} must be a CUDA tensor (there is no CPU fallback)")
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise TypeError(f"se3gnn_b200.message_layer: {name} must be contiguous float32")
    return MsgLayerFn.apply(xe, w1[0], w1[1], n1[0], n1[1], w2[0], w2[1], n2[0], n2[1], y, extra, ei, ns, nv, cs, cg, plan2)
