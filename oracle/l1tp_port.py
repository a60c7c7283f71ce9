"""TEST INFRASTRUCTURE ONLY — PyTorch CPU port of the reference tensor product.

``L1TPPort`` restates ``L1TensorProduct.forward`` of
/root/reference/models/segnn/l1_tensor_prod.py:234-299 with the same kind and order of eager
ATen ops the reference issues (boolean-mask gather, vecdot / cross, cat, matmul / tensordot,
masked write, in-place norm multiply), so that timing it on host cores is a fair stand-in for the
reference's own CPU path on machines where /root/reference is absent (the GPU box), and so that
autograd through it yields the gradients the CUDA backward is checked against.

Pinned by ``tests/test_oracle_port.py``: against the golden vectors from the unmodified
reference (always) and against the reference itself (when /root/reference is mounted).
Only tests/, smoke() and bench.py's cpu_baseline / --impl reference legs may import it.
"""
from __future__ import annotations

import math

import torch

from . import l1tp_oracle as O

C3 = 1.0 / math.sqrt(3.0)
C6 = 1.0 / math.sqrt(6.0)


class L1TPPort(torch.nn.Module):
    def __init__(self, in1: str, out: str, **norm_kwargs):
        super().__init__()
        self.in1_s, self.out_s = in1, out
        i1, io = O.parse_irreps(in1), O.parse_irreps(out)
        self.d_in, self.d_out = O.irreps_dim(i1), O.irreps_dim(io)
        ci, co = O.species_columns(i1), O.species_columns(io)
        # boolean masks like the reference (plain attributes, L1TP:24-65)
        self.mi, self.mo = {}, {}
        for sp in O.SPECIES:
            mi = torch.zeros(self.d_in, dtype=torch.bool)
            mo = torch.zeros(self.d_out, dtype=torch.bool)
            w = 1 if sp[0] == "0" else 3
            for c in ci[sp]:
                mi[c:c + w] = True
            for c in co[sp]:
                mo[c:c + w] = True
            self.mi[sp], self.mo[sp] = mi, mo
        self.n = {sp: len(ci[sp]) for sp in O.SPECIES}
        self.m = {sp: len(co[sp]) for sp in O.SPECIES}
        a, _, _ = O.norm_factors(i1, io, **norm_kwargs)
        for k, v in O.norm_buffers(io, a).items():
            self.register_buffer(k, torch.tensor(v, dtype=torch.float32))
        for k, shp in O.weight_shapes(i1, io).items():
            self.register_parameter(k, torch.nn.Parameter(torch.rand(shp) * 2 - 1))

    def forward(self, in1: torch.Tensor, in2: torch.Tensor) -> torch.Tensor:
        E = in1.shape[0]
        out = torch.zeros((E, self.d_out), dtype=in1.dtype)
        y0 = in2[:, 0:1]
        y1 = in2[:, None, 1:4]
        n, m = self.n, self.m
        for sp, ssp, vsp in (("0e", "0e", "1o"), ("0o", "0o", "1e")):
            if m[sp] == 0:
                continue
            parts = [in1[:, self.mi[ssp]] * y0]
            if n[vsp] > 0:
                parts.append(C3 * torch.linalg.vecdot(in1[:, self.mi[vsp]].reshape(E, n[vsp], 3), y1))
            res = torch.cat(parts, -1) @ getattr(self, f"weights_l{sp}")
            out[:, self.mo[sp]] = res.to(out.dtype)
            out[:, self.mo[sp]] *= getattr(self, f"norm_l{sp}")
        for sp, ssp, vsp, xsp in (("1e", "0o", "1e", "1o"), ("1o", "0e", "1o", "1e")):
            if m[sp] == 0:
                continue
            parts = [C3 * in1[:, self.mi[ssp], None] * y1]
            if n[vsp] > 0:
                parts.append(C3 * in1[:, self.mi[vsp]].reshape(E, n[vsp], 3) * in2[:, None, 0:1])
            if n[xsp] > 0:
                parts.append(C6 * torch.linalg.cross(in1[:, self.mi[xsp]].reshape(E, n[xsp], 3), y1))
            res = torch.tensordot(torch.cat(parts, -2), getattr(self, f"weights_l{sp}"), ([-2], [0]))
            out[:, self.mo[sp]] = res.transpose(-1, -2).reshape(E, 3 * m[sp])
            out[:, self.mo[sp]] *= getattr(self, f"norm_l{sp}")
        return out.contiguous()
