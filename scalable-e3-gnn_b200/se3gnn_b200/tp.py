"""Host side of the fused l<=1 tensor-product layer: species tables, plan cache and
the ``torch.autograd.Function`` that calls the CUDA kernels through the C ABI.

One call = one persistent kernel doing
    gather (virtual concat of indexed row segments) -> CG tensor product with SH(1)
    -> weight contraction -> norm -> [swish/sigmoid gate] -> [residual] -> [sorted-segment sum]
and one more kernel for the backward (plus a tiny deterministic weight-gradient reduction).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import capi
from .irreps import Irreps, as_irreps

SPECIES = ("0e", "0o", "1e", "1o")  # index order used by the C ABI (L1TP:24-27)


def species_columns(irreps: Irreps) -> List[List[int]]:
    """Flat column of every channel per species (x component for l=1), declaration order.

    Equivalent to the boolean masks of the reference (L1TP:24-36, 53-65): same-species
    blocks of interleaved irreps are concatenated in order."""
    cols: List[List[int]] = [[], [], [], []]
    i = 0
    for mi in irreps:
        l, p = mi.ir.l, mi.ir.p
        if l > 1:
            raise AssertionError("only l<=1 irreps are supported")
        s = (0 if p == 1 else 1) if l == 0 else (2 if p == 1 else 3)
        step = 2 * l + 1
        cols[s].extend(range(i, i + mi.mul * step, step))
        i += mi.mul * step
    return cols


_PLANS: Dict[Tuple[str, str, int], capi.L1tpPlan] = {}


def get_plan(in1: Irreps, out: Irreps) -> capi.L1tpPlan:
    dev = torch.cuda.current_device()
    key = (str(in1), str(out), dev)
    p = _PLANS.get(key)
    if p is None:
        ic, oc = species_columns(in1), species_columns(out)
        p = capi.L1tpPlan([len(c) for c in ic], [len(c) for c in oc], ic, oc)
        _PLANS[key] = p
    return p


@dataclass
class TPConfig:
    """Static (non-tensor) description of one fused TP-layer call."""
    plan: capi.L1tpPlan
    widths: Sequence[int]                 # width of each in1 segment
    epilogue: int = capi.EPI_RAW
    gate_ns: int = 0
    gate_cs: float = 1.0
    gate_cg: float = 1.0
    grad_modes: Sequence[int] = ()        # per segment; default derived from idx
    num_segments: int = 0                 # rows of the segment-sum output
    need_gin2: bool = False
    # which of the segments share one gradient buffer (e.g. x gathered by dst and by src)
    share_grad: Dict[int, int] = field(default_factory=dict)

    tag: str = "tp"

    @property
    def d_post(self) -> int:
        if self.epilogue == capi.EPI_GATE:
            return self.gate_ns + 3 * self.plan.m[3]
        return self.plan.d_out

    def flops_fwd_per_row(self) -> int:
        """Minimal (factorised) FMA count x2 of one forward row, see csrc/l1tp.cu header."""
        n, m = self.plan.n, self.plan.m
        fe = (n[0] + n[3]) * m[0] + n[0] * m[3] + 3 * (n[3] + n[2]) * m[3]
        fo = (n[1] + n[2]) * m[1] + n[1] * m[2] + 3 * (n[2] + n[3]) * m[2]
        return 2 * (fe + fo)

    def algo_bytes(self, rows: int, segs, idxs, backward: bool, nseg_out: int = 0, has_resid: bool = False,
                   grads=None, part: Optional[str] = None) -> float:
        """Algorithmic HBM bytes of one launch: every operand once; a segment gathered through a
        *sorted* index is charged once per distinct row (segment-cached), an unsorted gather once per row.
        ``part`` (backward only): "w" = weight-gradient kernel alone (reads the in1 segments, writes only the weight
        gradients), "i" = input-gradient kernel alone (does not read the in1 segments)."""
        b = rows * 4 * 4  # in2
        for i, s in enumerate(segs):
            if backward and part == "i":
                break
            w = self.widths[i]
            if idxs[i] is None:
                b += rows * w * 4
            else:
                srt = i < len(self.grad_modes) and self.grad_modes[i] == capi.GRAD_SORTED
                b += (min(rows, s.shape[0]) if srt else rows) * w * 4 + rows * 4
        gate = self.epilogue == capi.EPI_GATE
        d_out, d_post = self.plan.d_out, self.d_post
        if not backward:
            if nseg_out:
                b += nseg_out * d_post * 4 + rows * 4
                if gate:
                    b += rows * d_out * 4  # pre-activation saved for backward
            else:
                b += rows * d_out * 4 + (rows * d_post * 4 if gate else 0)
            if has_resid:
                b += rows * d_out * 4
        else:
            if gate:
                b += rows * d_out * 4
            b += (nseg_out if nseg_out else rows) * d_post * 4 + (rows * 4 if nseg_out else 0)
            for i, s in enumerate(segs):
                if part == "w":
                    break
                if grads is not None and grads[i]:
                    w = self.widths[i]
                    srt = idxs[i] is not None and i < len(self.grad_modes) and self.grad_modes[i] == capi.GRAD_SORTED
                    b += (min(rows, s.shape[0]) if srt else rows) * w * 4
        return float(b)


def _chk(t: Optional[torch.Tensor], name: str, dtype=torch.float32):
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError(f"se3gnn_b200: {name} must be a CUDA tensor (there is no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"se3gnn_b200: {name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"se3gnn_b200: {name} must be contiguous")


class TPLayerFn(torch.autograd.Function):
    """args: cfg, in2, resid, seg_idx, gout_rows(int), idx tensors (tuple), weights(4), norms(4), *segments"""

    @staticmethod
    def forward(ctx, cfg: TPConfig, rows: int, in2, resid, seg_idx, idxs, norms, w0, w1, w2, w3, *segs):
        lib = capi.lib()
        ws = (w0, w1, w2, w3)
        nseg = len(segs)
        assert nseg == len(cfg.widths) == len(idxs)
        _chk(in2, "in2")
        _chk(resid, "resid")
        _chk(seg_idx, "seg_idx", torch.int32)
        for i, s in enumerate(segs):
            _chk(s, f"segment {i}")
            _chk(idxs[i], f"idx {i}", torch.int32)
        for i, w in enumerate(ws):
            _chk(w, f"weight {i}")
            _chk(norms[i], f"norm {i}")
        dev = in2.device
        a = capi.L1tpFwdArgs()
        a.rows = rows
        a.nseg = nseg
        for i, s in enumerate(segs):
            a.seg[i].base = s.data_ptr()
            a.seg[i].idx = capi.ptr(idxs[i])
            a.seg[i].width = int(cfg.widths[i])
            a.seg[i].ld = int(s.shape[-1])
        a.in2 = in2.data_ptr()
        for i in range(4):
            a.w[i] = capi.ptr(ws[i])
            a.norm[i] = capi.ptr(norms[i])
        a.epilogue = cfg.epilogue
        a.gate_ns = cfg.gate_ns
        a.gate_cs = cfg.gate_cs
        a.gate_cg = cfg.gate_cg
        d_out, d_post = cfg.plan.d_out, cfg.d_post
        gate = cfg.epilogue == capi.EPI_GATE
        need_raw = gate or seg_idx is None
        raw = torch.empty((rows, d_out), device=dev, dtype=torch.float32) if need_raw else None
        post = None
        out_seg = None
        if seg_idx is not None:
            out_seg = torch.zeros((cfg.num_segments, d_post), device=dev, dtype=torch.float32)
            a.seg_idx = seg_idx.data_ptr()
            a.out_seg = out_seg.data_ptr()
        elif gate:
            post = torch.empty((rows, d_post), device=dev, dtype=torch.float32)
            a.out_post = post.data_ptr()
        if raw is not None:
            a.out_raw = raw.data_ptr()
        a.resid = capi.ptr(resid)
        with capi.mark(cfg.tag + ".fwd",
                       cfg.algo_bytes(rows, segs, idxs, False, cfg.num_segments if seg_idx is not None else 0,
                                      resid is not None) if capi._prof is not None else 0.0,
                       float(cfg.flops_fwd_per_row()) * rows):
            capi.check(lib.se3_l1tp_forward(cfg.plan.handle, C.byref(a), capi.current_stream_ptr()), "se3_l1tp_forward")
        ctx.cfg = cfg
        ctx.rows = rows
        ctx.idxs = idxs
        ctx.norms = norms
        ctx.seg_idx = seg_idx
        ctx.has_resid = resid is not None
        ctx.save_for_backward(in2, raw if gate else None, *ws, *segs)
        if out_seg is not None:
            return out_seg
        return post if gate else raw

    @staticmethod
    def backward(ctx, gout):
        lib = capi.lib()
        cfg: TPConfig = ctx.cfg
        saved = ctx.saved_tensors
        in2, raw = saved[0], saved[1]
        ws = saved[2:6]
        segs = saved[6:]
        nseg = len(segs)
        gout = gout.contiguous()
        _chk(gout, "grad_output")
        dev = in2.device
        a = capi.L1tpBwdArgs()
        a.rows = ctx.rows
        a.nseg = nseg
        for i, s in enumerate(segs):
            a.seg[i].base = s.data_ptr()
            a.seg[i].idx = capi.ptr(ctx.idxs[i])
            a.seg[i].width = int(cfg.widths[i])
            a.seg[i].ld = int(s.shape[-1])
        a.in2 = in2.data_ptr()
        for i in range(4):
            a.w[i] = capi.ptr(ws[i])
            a.norm[i] = capi.ptr(ctx.norms[i])
        a.epilogue = cfg.epilogue
        a.gate_ns = cfg.gate_ns
        a.gate_cs = cfg.gate_cs
        a.gate_cg = cfg.gate_cg
        a.raw = capi.ptr(raw)
        a.gout = gout.data_ptr()
        a.gout_idx = capi.ptr(ctx.seg_idx)
        # needs_input_grad: (cfg, rows, in2, resid, seg_idx, idxs, norms, w0..w3, *segs)
        nig = ctx.needs_input_grad
        seg_need = nig[11:11 + nseg]
        w_need = nig[7:11]
        gsegs: List[Optional[torch.Tensor]] = [None] * nseg
        for i, s in enumerate(segs):
            if not seg_need[i]:
                continue
            owner = cfg.share_grad.get(i, i)
            idx = ctx.idxs[i]
            if owner != i and gsegs[owner] is not None:
                buf = gsegs[owner]
            elif idx is None:
                buf = torch.empty_like(s)
                if s.shape[-1] != cfg.widths[i] or s.shape[0] != ctx.rows:
                    buf.zero_()
            else:
                buf = torch.zeros_like(s)
            gsegs[i] = buf
            a.gseg[i] = buf.data_ptr()
            mode = cfg.grad_modes[i] if i < len(cfg.grad_modes) and cfg.grad_modes[i] else (
                capi.GRAD_STORE if idx is None else capi.GRAD_ATOMIC)
            a.gseg_mode[i] = mode
        gws: List[Optional[torch.Tensor]] = [None] * 4
        if any(w_need):
            for i in range(4):
                if ws[i] is not None:
                    gws[i] = torch.empty_like(ws[i])
                    a.gw[i] = gws[i].data_ptr()
        gin2 = None
        if nig[2] and cfg.need_gin2:
            gin2 = torch.empty_like(in2)
            a.gin2 = gin2.data_ptr()
        nso = cfg.num_segments if ctx.seg_idx is not None else 0
        if capi._prof is None:
            capi.check(lib.se3_l1tp_backward(cfg.plan.handle, C.byref(a), capi.current_stream_ptr()), "se3_l1tp_backward")
        else:
            # per-kernel table of bench.py: the C ABI runs the weight-gradient and the input-gradient kernels of one call
            # back to back; time them separately by asking for one output group per call (same kernels, same results)
            want_w, want_i = any(g is not None for g in gws), any(g is not None for g in gsegs)
            keep_w, keep_i = [a.gw[i] for i in range(4)], [a.gseg[i] for i in range(capi.MAX_SEG)]
            if want_w:
                keep_gin2 = a.gin2
                a.gin2 = None
                for i in range(capi.MAX_SEG):
                    a.gseg[i] = None
                with capi.mark(cfg.tag + ".bwdw", cfg.algo_bytes(ctx.rows, segs, ctx.idxs, True, nso, False, seg_need, "w"),
                               float(cfg.flops_fwd_per_row()) * ctx.rows):
                    capi.check(lib.se3_l1tp_backward(cfg.plan.handle, C.byref(a), capi.current_stream_ptr()), "se3_l1tp_backward")
                for i in range(capi.MAX_SEG):
                    a.gseg[i] = keep_i[i]
                a.gin2 = keep_gin2
            if want_i or gin2 is not None:
                for i in range(4):
                    a.gw[i] = None
                with capi.mark(cfg.tag + ".bwdi", cfg.algo_bytes(ctx.rows, segs, ctx.idxs, True, nso, False, seg_need, "i"),
                               float(cfg.flops_fwd_per_row()) * ctx.rows):
                    capi.check(lib.se3_l1tp_backward(cfg.plan.handle, C.byref(a), capi.current_stream_ptr()), "se3_l1tp_backward")
                for i in range(4):
                    a.gw[i] = keep_w[i]
        gresid = gout if (ctx.has_resid and nig[3]) else None
        out_gsegs = []
        for i in range(nseg):
            owner = cfg.share_grad.get(i, i)
            out_gsegs.append(gsegs[i] if (owner == i or gsegs[owner] is None or not seg_need[owner]) else None)
        return (None, None, gin2, gresid, None, None, None, *[g if n else None for g, n in zip(gws, w_need)],
                *out_gsegs)


def tp_layer(cfg: TPConfig, rows: int, segs: Sequence[torch.Tensor], idxs: Sequence[Optional[torch.Tensor]],
             in2: torch.Tensor, weights: Sequence[Optional[torch.Tensor]], norms: Sequence[Optional[torch.Tensor]],
             resid: Optional[torch.Tensor] = None, seg_idx: Optional[torch.Tensor] = None) -> torch.Tensor:
    return TPLayerFn.apply(cfg, rows, in2, resid, seg_idx, tuple(idxs), tuple(norms), *weights, *segs)
