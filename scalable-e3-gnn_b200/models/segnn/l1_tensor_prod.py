"""``L1TensorProduct`` — drop-in for the reference module of the same name and import path
(``/root/reference/models/segnn/l1_tensor_prod.py:8-299``, cited as ``L1TP:<line>``).

Same constructor, same public attributes, same ``state_dict`` keys/shapes, same
initialisation stream (an identical ``torch.manual_seed`` gives identical weights) and the
same error behaviour; ``forward`` runs one fused sm_100a kernel (and one for backward)
through the C ABI in ``include/se3gnn_b200.h`` instead of ~40 eager ATen ops.

There is no CPU path: ``forward`` on non-CUDA tensors raises.

Documented deviations from the reference:
* Output columns of zero-width species cannot exist, so nothing is left uninitialised
  (the reference returns ``torch.empty`` memory there, L1TP:240).
* Arithmetic is fp32.  Inputs / parameters of another float dtype are converted to fp32 on
  the device and the result is converted back, so the bf16-autocast failure of the
  reference (SURVEY Q5, L1TP:281,295) does not occur.
"""
from __future__ import annotations

from math import sqrt
from typing import List, Optional

import torch
from torch import Tensor
from torch.nn import Module

from se3gnn_b200 import capi
from se3gnn_b200.irreps import Instruction, Irreps, as_irreps
from se3gnn_b200.tp import TPConfig, get_plan, species_columns, tp_layer

_SP = ("l0e", "l0o", "l1e", "l1o")


def _species_masks(irreps: Irreps):
    """Boolean masks over the flat axis per species (L1TP:24-36)."""
    masks = {s: torch.zeros(irreps.dim, dtype=torch.bool) for s in _SP}
    pos = 0
    for mi in irreps:
        key = f"l{mi.ir.l}{'e' if mi.ir.p == 1 else 'o'}"
        if key in masks:
            masks[key][pos:pos + mi.dim] = True
        pos += mi.dim
    return masks


class L1TensorProduct(Module):
    def __init__(self, in1_irreps, out_irreps=None,
                 irrep_normalization="component", path_normalization="element",
                 in1_var: Optional[List[float]] = None, in2_var: Optional[List[float]] = None,
                 out_var: Optional[List[float]] = None) -> None:
        super().__init__()
        in1_irreps = as_irreps(in1_irreps)
        assert in1_irreps.lmax == 1                                   # L1TP:13
        if out_irreps is not None:
            out_irreps = as_irreps(out_irreps)
            assert out_irreps.lmax == 1                               # L1TP:14

        self.iri1 = in1_irreps
        self.iri2 = Irreps.spherical_harmonics(1)                     # 1x0e+1x1o, L1TP:17
        self.iro = out_irreps if out_irreps is not None else in1_irreps
        self.in1_dim = self.iri1.dim
        self.in2_dim = self.iri2.dim

        m1, m2, mo = _species_masks(self.iri1), _species_masks(self.iri2), _species_masks(self.iro)
        for s in _SP:
            setattr(self, f"iri1_{s}", m1[s])
            setattr(self, f"iro_{s}", mo[s])
        self.iri2_l0e, self.iri2_l1o = m2["l0e"], m2["l1o"]

        # species counts, attribute names as in L1TP:67-77
        self.num_i1_l0e = int(m1["l0e"].sum())
        self.num_i1_l0o = int(m1["l0o"].sum())
        self.num_i1_l0 = self.num_i1_l0e + self.num_i1_l0o
        self.dim_i1_l1e = int(m1["l1e"].sum())
        self.num_i1_l1e = self.dim_i1_l1e // 3
        self.dim_i1_l1o = int(m1["l1o"].sum())
        self.num_i1_l1o = self.dim_i1_l1o // 3
        self.dim_o_l0e = int(mo["l0e"].sum())
        self.dim_o_l0o = int(mo["l0o"].sum())
        self.dim_o_l1e = int(mo["l1e"].sum())
        self.dim_o_l1o = int(mo["l1o"].sum())

        # Weight rows follow the feature order of the forward (L1TP:81-88).  The RNG calls are
        # issued in the reference's order so that a given seed reproduces its initial weights.
        n0e, n0o, n1e, n1o = self.num_i1_l0e, self.num_i1_l0o, self.num_i1_l1e, self.num_i1_l1o
        rows = {"l0e": n0e + n1o, "l0o": n0o + n1e, "l1e": n0o + n1e + n1o, "l1o": n0e + n1o + n1e}
        cols = {"l0e": self.dim_o_l0e, "l0o": self.dim_o_l0o, "l1e": self.dim_o_l1e // 3, "l1o": self.dim_o_l1o // 3}
        for s in _SP:
            if rows[s] > 0 and cols[s] > 0:
                setattr(self, f"weights_{s}", torch.nn.Parameter(torch.rand((rows[s], cols[s])) * 2 - 1))

        self.cg000 = 1
        self.cg110 = 1 / sqrt(3)
        self.cg011 = self.cg110
        self.cg111 = 1 / sqrt(6)

        def _vars(v, irreps, msg):
            if v is None:
                return [1.0] * len(irreps)
            v = [float(x) for x in v]
            assert len(v) == len(irreps), msg
            return v

        in1_var = _vars(in1_var, self.iri1, "Len of ir1_var must be equal to len(irreps_in1)")
        in2_var = _vars(in2_var, self.iri2, "Len of ir2_var must be equal to len(irreps_in2)")
        out_var = _vars(out_var, self.iro, "Len of out_var must be equal to len(irreps_out)")

        self.is_norm = irrep_normalization in ("component", "norm") or path_normalization in ("element", "path")
        if not self.is_norm:
            return  # reference quirk Q3 (L1TP:116): is_comp_norm / instructions stay undefined
        self.is_comp_norm = irrep_normalization != "norm" and path_normalization != "path"
        torch._assert(self.is_comp_norm, "Not all norms are implemented yet.")

        self.instructions: List[Instruction] = []
        for s in _SP:
            n = {"l0e": self.dim_o_l0e, "l0o": self.dim_o_l0o, "l1e": self.dim_o_l1e, "l1o": self.dim_o_l1o}[s]
            self.register_buffer(f"norm_{s}", torch.empty(n))
        # One cursor per species, shared by the norm buffer and the weight-column re-init exactly as
        # in L1TP:163-189.  For l=1 it advances by 3*mul, so when an l=1 species appears in several
        # output irreps the later blocks keep their first U(-1,1) draw (reference quirk, kept so that
        # a given seed reproduces the reference's initial weights bit for bit).
        fill = {s: 0 for s in _SP}
        for io, mir_out in enumerate(self.iro):
            lo, po = mir_out.ir.l, mir_out.ir.p
            alpha = mir_out.ir.dim * out_var[io] if irrep_normalization == "component" else 1
            x = 0.0
            first = len(self.instructions)
            for ii2, mir_in2 in enumerate(self.iri2):
                for ii1, mir_in1 in enumerate(self.iri1):
                    l1, l2 = mir_in1.ir.l, mir_in2.ir.l
                    # Reference precedence (L1TP:137-138): `A or (B and C)`, so parity is NOT
                    # checked for l=0 outputs.
                    if (lo == 0 and l2 == l1) or (lo == 1 and (l2 | l1) and po == mir_in2.ir.p * mir_in1.ir.p):
                        x += in1_var[ii1] * in2_var[ii2] * mir_in1.mul * mir_in2.mul
                        self.instructions.append(
                            Instruction(ii1, ii2, io, "uvw", True, alpha, (mir_in1.mul, mir_in2.mul, mir_out.mul)))
            if path_normalization == "none":
                a, wi = sqrt(alpha), 1 / sqrt(x)
            else:
                a, wi = sqrt((alpha / x) if x > 0 else alpha), 1
            s = f"l{lo}{'e' if po == 1 else 'o'}"
            with torch.no_grad():
                getattr(self, f"norm_{s}")[fill[s]:fill[s] + mir_out.dim] = a
                getattr(self, f"weights_{s}")[:, fill[s]:fill[s] + mir_out.mul].uniform_(-wi, wi)
            fill[s] += mir_out.dim
            for i in range(first, len(self.instructions)):
                ins = self.instructions[i]
                self.instructions[i] = Instruction(*ins[:-2], path_weight=a, path_shape=ins.path_shape)

    # ------------------------------------------------------------------ CUDA path
    def _cfg(self, need_gin2: bool) -> TPConfig:
        # the plan (a ctypes handle with device tables) is looked up per call in the (irreps, device) cache and never
        # becomes module state: copy.deepcopy / pickle / torch.save of the module work as they do for the reference
        return TPConfig(plan=get_plan(self.iri1, self.iro), widths=(self.in1_dim,), need_gin2=need_gin2)

    def forward(self, in1: Tensor, in2: Tensor) -> Tensor:
        torch._assert(in1.shape[-1] == self.in1_dim,
                      f"Incorrect last dimension for in1 = {in1.shape[-1]}, required is {self.in1_dim}")
        torch._assert(in2.shape[-1] == self.in2_dim,
                      f"Incorrect last dimension for in2 = {in2.shape[-1]}, required is {self.in2_dim}")
        comp_norm = self.is_comp_norm  # AttributeError when both normalisations are "none" (quirk Q3)
        if in1.dim() != 2 or in2.dim() != 2:
            raise IndexError("L1TensorProduct expects 2-D [rows, dim] inputs (reference quirk Q4)")
        if in2.shape[0] != in1.shape[0]:
            raise RuntimeError(f"in1 has {in1.shape[0]} rows, in2 has {in2.shape[0]}")
        if not in1.is_cuda:
            raise RuntimeError("se3gnn_b200.L1TensorProduct runs on CUDA (sm_100a) only; there is no CPU fallback")
        dt = in1.dtype
        x = in1.to(torch.float32).contiguous()
        y = in2.to(device=in1.device, dtype=torch.float32).contiguous()
        ws, norms = [], []
        for s in _SP:
            w = getattr(self, f"weights_{s}", None)
            ws.append(None if w is None else w.to(torch.float32).contiguous())
            nb = getattr(self, f"norm_{s}", None) if comp_norm else None
            norms.append(None if nb is None or nb.numel() == 0 else nb.to(torch.float32).contiguous())
        cfg = self._cfg(need_gin2=bool(y.requires_grad))
        out = tp_layer(cfg, x.shape[0], [x], [None], y, ws, norms)
        return out.to(dt).contiguous()
