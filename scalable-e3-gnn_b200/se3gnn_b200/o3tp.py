"""Fully connected O(3) tensor product for l <= 2 with a spherical-harmonics type second input: the generalisation of
the reference's ``L1TensorProduct`` that BASELINE configs[2] (SEGNN l_max = 2) needs.  The reference excludes l = 2
(``/root/reference/models/segnn/l1_tensor_prod.py:13-14``); this module keeps its conventions — constructor argument
order, ``iri1/iri2/iro/in1_dim/in2_dim/instructions`` attributes (``L1TP:16-21,121,151``), path enumeration
``(i_out, i_in2, i_in1)``, 'component' x 'element' normalisation (``L1TP:124,145,169``) — so that for l <= 1 irreps of
SH type it computes exactly what ``L1TensorProduct`` does (weights laid out per path instead of stacked per species).

All arithmetic runs in the CUDA library (``se3_o3tp_forward/backward``, csrc/o3tp.cu); there is no eager fallback.
"""
from __future__ import annotations

import torch

from . import capi
from .irreps import Instruction, Irreps, as_irreps

__all__ = ["O3TensorProduct"]


class _O3tpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, in1, in2, weight, mod):
        plan = mod._plan
        rows = in1.shape[0]
        out = torch.empty((rows, plan.d_out), device=in1.device, dtype=torch.float32)
        with capi.mark(f"o3tp.fwd[{mod.in1_dim}x{mod.in2_dim}->{mod.iro.dim}]", mod.algo_bytes(rows, "fwd"), mod.flops(rows)):
            capi.check(capi.lib().se3_o3tp_forward(plan.handle, rows, capi.ptr(in1), capi.ptr(in2), capi.ptr(weight),
                                                   capi.ptr(out), capi.current_stream_ptr()), "se3_o3tp_forward")
        ctx.save_for_backward(in1, in2, weight)
        ctx.mod = mod
        return out

    @staticmethod
    def backward(ctx, gout):
        in1, in2, weight = ctx.saved_tensors
        mod = ctx.mod
        plan = mod._plan
        rows = in1.shape[0]
        gout = gout.contiguous()
        gin1 = torch.empty_like(in1)
        gin2 = torch.empty_like(in2) if ctx.needs_input_grad[1] else None
        gw = torch.empty_like(weight)
        with capi.mark(f"o3tp.bwd[{mod.in1_dim}x{mod.in2_dim}->{mod.iro.dim}]", mod.algo_bytes(rows, "bwd"), 2 * mod.flops(rows)):
            capi.check(capi.lib().se3_o3tp_backward(plan.handle, rows, capi.ptr(in1), capi.ptr(in2), capi.ptr(weight),
                                                    capi.ptr(gout), capi.ptr(gin1), capi.ptr(gin2), capi.ptr(gw),
                                                    capi.current_stream_ptr()), "se3_o3tp_backward")
        return gin1, gin2, gw, None


def _rowsegs(tensors, idxs):
    segs = (capi.RowSeg * capi.MAX_SEG)()
    for s, (t, ix) in enumerate(zip(tensors, idxs)):
        segs[s].base, segs[s].idx = capi.ptr(t), capi.ptr(ix)
        segs[s].width, segs[s].ld = t.shape[1], t.stride(0)
    return segs


class _O3tpCatFn(torch.autograd.Function):
    """in1 = cat_s(tensors[s][idxs[s]] if idxs[s] is not None else tensors[s]) read in place by the kernels."""

    @staticmethod
    def forward(ctx, in2, weight, mod, idxs, sorted_flags, *tensors):
        plan = mod._plan
        rows = in2.shape[0]
        out = torch.empty((rows, plan.d_out), device=in2.device, dtype=torch.float32)
        with capi.mark(f"o3tp.fwd[{mod.in1_dim}x{mod.in2_dim}->{mod.iro.dim}]", mod.algo_bytes(rows, "fwd"), mod.flops(rows)):
            capi.check(capi.lib().se3_o3tp_forward_seg(plan.handle, rows, len(tensors), _rowsegs(tensors, idxs), capi.ptr(in2),
                                                       capi.ptr(weight), capi.ptr(out), capi.current_stream_ptr()),
                       "se3_o3tp_forward_seg")
        ctx.save_for_backward(in2, weight, *tensors)
        ctx.mod, ctx.idxs, ctx.sorted_flags = mod, idxs, sorted_flags
        return out

    @staticmethod
    def backward(ctx, gout):
        in2, weight, *tensors = ctx.saved_tensors
        mod, idxs, sorted_flags = ctx.mod, ctx.idxs, ctx.sorted_flags
        plan = mod._plan
        rows = in2.shape[0]
        gout = gout.contiguous()
        grads, gptr, modes = [], (capi.C.c_void_p * capi.MAX_SEG)(), (capi.C.c_int32 * capi.MAX_SEG)()
        for s, (t, ix) in enumerate(zip(tensors, idxs)):
            if not ctx.needs_input_grad[5 + s]:
                grads.append(None)
                modes[s] = capi.GRAD_NONE
                continue
            # same row stride as the source: the kernels address source and gradient alike
            g = torch.empty_like(t) if ix is None else torch.zeros_like(t)
            if g.stride() != t.stride():
                raise capi.Se3Error("o3tp: segment tensors need a dense row-major layout")
            grads.append(g)
            gptr[s] = capi.ptr(g)
            modes[s] = capi.GRAD_STORE if ix is None else (capi.GRAD_SORTED if sorted_flags[s] else capi.GRAD_ATOMIC)
        gin2 = torch.empty_like(in2) if ctx.needs_input_grad[0] else None
        gw = torch.empty_like(weight)
        with capi.mark(f"o3tp.bwd[{mod.in1_dim}x{mod.in2_dim}->{mod.iro.dim}]", mod.algo_bytes(rows, "bwd"), 2 * mod.flops(rows)):
            capi.check(capi.lib().se3_o3tp_backward_seg(plan.handle, rows, len(tensors), _rowsegs(tensors, idxs),
                                                        capi.ptr(in2), capi.ptr(weight), capi.ptr(gout), gptr, modes,
                                                        capi.ptr(gin2), capi.ptr(gw), capi.current_stream_ptr()),
                       "se3_o3tp_backward_seg")
        return (gin2, gw, None, None, None, *grads)


_PLANS: dict = {}


class O3TensorProduct(torch.nn.Module):
    def __init__(self, in1_irreps, out_irreps=None, in2_irreps=None):
        super().__init__()
        self.iri1 = as_irreps(in1_irreps)
        self.iri2 = Irreps.spherical_harmonics(2) if in2_irreps is None else as_irreps(in2_irreps)
        self.iro = self.iri1 if out_irreps is None else as_irreps(out_irreps)
        assert max(self.iri1.lmax, self.iri2.lmax, self.iro.lmax) <= 2, "Maximal l supported by this tensor product is 2."
        assert all(mi.mul == 1 for mi in self.iri2), "in2 must have multiplicity 1 per irrep (spherical harmonics type)"
        self.in1_dim, self.in2_dim = self.iri1.dim, self.iri2.dim
        _ = self._plan   # created now: configuration errors surface at construction
        assert (self._plan.d_in1, self._plan.d_in2, self._plan.d_out) == (self.in1_dim, self.in2_dim, self.iro.dim)
        self.instructions = [
            Instruction(i1, i2, io, "uvw", True, a, (self.iri1[i1].mul, 1, self.iro[io].mul))
            for i1, i2, io, _, a in self._plan.paths
        ]
        self.weight_offsets = [p[3] for p in self._plan.paths]
        self.weight = torch.nn.Parameter(torch.randn(self._plan.weight_floats))

    @property
    def _plan(self) -> "capi.O3tpPlan":
        """The plan (ctypes handle + device tables on the CURRENT device) lives in a per-(irreps, device) cache, not in
        the module: deepcopy / pickle / torch.save of the module work, and ``.to(other_gpu)`` gets tables there."""
        key = (str(self.iri1), str(self.iri2), str(self.iro), torch.cuda.current_device() if torch.cuda.is_available() else -1)
        p = _PLANS.get(key)
        if p is None:
            p = _PLANS[key] = capi.O3tpPlan([(mi.mul, mi.ir.l, mi.ir.p) for mi in self.iri1],
                                            [(mi.ir.l, mi.ir.p) for mi in self.iri2],
                                            [(mi.mul, mi.ir.l, mi.ir.p) for mi in self.iro])
        return p

    def weight_views(self):
        """Per-path [mul_in1, mul_out] views of the flat weight, in `instructions` order."""
        return [self.weight[o:o + ins.path_shape[0] * ins.path_shape[2]].view(ins.path_shape[0], ins.path_shape[2])
                for o, ins in zip(self.weight_offsets, self.instructions)]

    def flops(self, rows: int) -> float:
        """Algorithmic flops of the weight contraction, forward (2 per multiply-add)."""
        return float(rows) * sum(2.0 * ins.path_shape[0] * ins.path_shape[2] * self.iro[ins.i_out].ir.dim
                                 for ins in self.instructions)

    def algo_bytes(self, rows: int, part: str) -> float:
        """Algorithmic HBM bytes (the TP-standalone formula): every operand once, fp32."""
        d1, d2, do = self.in1_dim, self.in2_dim, self.iro.dim
        per_row = (d1 + d2 + do) if part == "fwd" else (do + d1 + d2 + d1 + d2)
        return 4.0 * (rows * per_row + self._plan.weight_floats)

    def forward_cat(self, parts, in2: torch.Tensor) -> torch.Tensor:
        """TP(cat(parts), in2) without materialising the concatenation: parts = [(tensor [n_s, width_s], idx or None[,
        sorted])] (<= 4), idx an int32 [rows] gather index into the tensor's rows (e.g. the edge list), None = one row per
        output row.  The gradient of a gathered part is scatter-added inside the backward kernel; `sorted=True` promises a
        sorted index (the dst list of the graph), whose runs are summed on chip before one atomic add per run."""
        tensors, idxs, flags = [], [], []
        for part in parts:
            t, ix = part[0], part[1]
            flags.append(bool(part[2]) if len(part) > 2 else False)
            if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2):
                raise capi.Se3Error("O3TensorProduct.forward_cat: parts must be CUDA fp32 matrices")
            if t.stride(1) != 1:
                t = t.contiguous()
            if ix is not None:
                if ix.dtype != torch.int32 or ix.shape[0] != in2.shape[0]:
                    raise capi.Se3Error("O3TensorProduct.forward_cat: gather indices must be int32 [rows]")
                ix = ix.contiguous()
            elif t.shape[0] != in2.shape[0]:
                raise capi.Se3Error("O3TensorProduct.forward_cat: an ungathered part needs one row per output row")
            tensors.append(t)
            idxs.append(ix)
        torch._assert(sum(t.shape[1] for t in tensors) == self.in1_dim, "Incorrect last dimension for in1")
        torch._assert(in2.dim() == 2 and in2.shape[-1] == self.in2_dim, "Incorrect last dimension for in2")
        if len(tensors) > capi.MAX_SEG:
            raise capi.Se3Error(f"O3TensorProduct.forward_cat: at most {capi.MAX_SEG} parts")
        return _O3tpCatFn.apply(in2.contiguous(), self.weight, self, tuple(idxs), tuple(flags), *tensors)

    def forward(self, in1: torch.Tensor, in2: torch.Tensor) -> torch.Tensor:
        torch._assert(in1.dim() == 2 and in1.shape[-1] == self.in1_dim, "Incorrect last dimension for in1")
        torch._assert(in2.dim() == 2 and in2.shape[-1] == self.in2_dim, "Incorrect last dimension for in2")
        torch._assert(in1.shape[0] == in2.shape[0], "in1 and in2 need the same number of rows")
        if not (in1.is_cuda and in2.is_cuda and self.weight.is_cuda):
            raise capi.Se3Error("O3TensorProduct runs on CUDA tensors only (no CPU fallback)")
        if in1.dtype != torch.float32 or in2.dtype != torch.float32:
            raise capi.Se3Error("O3TensorProduct computes in fp32; cast the inputs")
        return _O3tpFn.apply(in1.contiguous(), in2.contiguous(), self.weight, self)
