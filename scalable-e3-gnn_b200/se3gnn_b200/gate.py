"""Gate constants.  e3nn's ``Gate`` wraps its activations in ``normalize2mom`` so that
E_{z~N(0,1)}[act(z)^2] = 1.  e3nn estimates the constant from 1e6 random samples; here it
is computed by Gauss-Hermite quadrature (deterministic, ~1e-12 accurate).  [public-SEGNN:
O3TensorProductSwishGate = TP -> Gate(scalars: silu, gates: sigmoid)]."""
from __future__ import annotations

import math

import numpy as np


def normalize2mom_const(fn) -> float:
    x, w = np.polynomial.hermite.hermgauss(256)
    z = math.sqrt(2.0) * x
    m2 = float((w * fn(z) ** 2).sum() / math.sqrt(math.pi))
    return m2 ** -0.5


def _sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


SILU_CST = normalize2mom_const(lambda z: z * _sigmoid(z))
SIGMOID_CST = normalize2mom_const(_sigmoid)


_GATE_FNS: dict = {}


def _layout(ns, blocks):
    from . import capi
    blocks = tuple((int(c), int(d)) for c, d in blocks if int(c) > 0)
    cnt = (capi.C.c_int32 * 4)(*[c for c, _ in blocks])
    dim = (capi.C.c_int32 * 4)(*[d for _, d in blocks])
    d_out = ns + sum(c * d for c, d in blocks)
    d_raw = d_out + sum(c for c, _ in blocks)
    return blocks, cnt, dim, d_out, d_raw


def irreps_gate(raw, ns: int, blocks):
    """Gate on the CUDA kernels (`se3_gate_forward/backward`, csrc/gate.cu): raw [rows, ns + ng + sum(cnt*dim)] ->
    [rows, ns + sum(cnt*dim)] with blocks = [(cnt, dim), ...] (<= 4) and ng = sum(cnt).  Differentiable."""
    import torch

    from . import capi

    blocks, cnt, dim, d_out, d_raw = _layout(ns, blocks)
    key = ("gate", ns, blocks)
    fn = _GATE_FNS.get(key)
    if fn is not None:
        return fn.apply(raw)

    class _Gate(torch.autograd.Function):
        @staticmethod
        def forward(ctx, raw):
            raw = raw.contiguous()
            if not raw.is_cuda or raw.dtype != torch.float32 or raw.dim() != 2 or raw.shape[1] != d_raw:
                raise capi.Se3Error(f"irreps_gate: need a CUDA fp32 [rows, {d_raw}] tensor")
            out = torch.empty((raw.shape[0], d_out), device=raw.device, dtype=torch.float32)
            with capi.mark("gate.fwd", 4.0 * raw.shape[0] * (d_raw + d_out)):
                capi.check(capi.lib().se3_gate_forward(raw.shape[0], ns, len(blocks), cnt, dim, SILU_CST, SIGMOID_CST,
                                                       capi.ptr(raw), capi.ptr(out), capi.current_stream_ptr()),
                           "se3_gate_forward")
            ctx.save_for_backward(raw)
            return out

        @staticmethod
        def backward(ctx, gout):
            (raw,) = ctx.saved_tensors
            gout = gout.contiguous()
            graw = torch.empty_like(raw)
            with capi.mark("gate.bwd", 4.0 * raw.shape[0] * (2 * d_raw + d_out)):
                capi.check(capi.lib().se3_gate_backward(raw.shape[0], ns, len(blocks), cnt, dim, SILU_CST, SIGMOID_CST,
                                                        capi.ptr(raw), capi.ptr(gout), capi.ptr(graw),
                                                        capi.current_stream_ptr()), "se3_gate_backward")
            return graw

    _GATE_FNS[key] = _Gate
    return _Gate.apply(raw)


def irreps_gate_segment_sum(raw, ns: int, blocks, dst, rowptr, n_seg: int):
    """sum over the incoming edges of gate(raw): raw [E, d_raw] with the edges sorted by destination `dst` (int32 [E]),
    `rowptr` int64 [n_seg + 1] their CSR offsets -> [n_seg, d_out].  One kernel each way (`se3_gate_segment_sum_*`,
    csrc/gate.cu): the gated per-edge tensor is never written and the aggregation uses no atomics.  Replaces
    ``zeros.index_add_(0, dst, gate(raw))`` (public SEGNN: message 2's gate followed by the add-aggregation)."""
    import torch

    from . import capi

    blocks, cnt, dim, d_out, d_raw = _layout(ns, blocks)
    key = ("segsum", ns, blocks)
    fn = _GATE_FNS.get(key)
    if fn is not None:
        return fn.apply(raw, dst, rowptr, n_seg)

    class _GateSum(torch.autograd.Function):
        @staticmethod
        def forward(ctx, raw, dst, rowptr, n_seg):
            raw = raw.contiguous()
            if not raw.is_cuda or raw.dtype != torch.float32 or raw.dim() != 2 or raw.shape[1] != d_raw:
                raise capi.Se3Error(f"irreps_gate_segment_sum: need a CUDA fp32 [rows, {d_raw}] tensor")
            out = torch.empty((n_seg, d_out), device=raw.device, dtype=torch.float32)
            with capi.mark("gate.segsum_fwd", 4.0 * (raw.shape[0] * d_raw + n_seg * d_out)):
                capi.check(capi.lib().se3_gate_segment_sum_forward(raw.shape[0], capi.ptr(dst), n_seg, ns, len(blocks), cnt, dim,
                                                                   SILU_CST, SIGMOID_CST, capi.ptr(raw), capi.ptr(out),
                                                                   capi.current_stream_ptr()), "se3_gate_segment_sum_forward")
            ctx.save_for_backward(raw, dst)
            return out

        @staticmethod
        def backward(ctx, gout):
            raw, dst = ctx.saved_tensors
            gout = gout.contiguous()
            graw = torch.empty_like(raw)
            with capi.mark("gate.segsum_bwd", 4.0 * (raw.shape[0] * 2 * d_raw + gout.shape[0] * d_out)):
                capi.check(capi.lib().se3_gate_segment_sum_backward(raw.shape[0], capi.ptr(dst), ns, len(blocks), cnt, dim,
                                                                    SILU_CST, SIGMOID_CST, capi.ptr(raw), capi.ptr(gout),
                                                                    capi.ptr(graw), capi.current_stream_ptr()),
                           "se3_gate_segment_sum_backward")
            return graw, None, None, None

    _GATE_FNS[key] = _GateSum
    return _GateSum.apply(raw, dst, rowptr, n_seg)
