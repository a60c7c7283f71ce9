"""TEST INFRASTRUCTURE ONLY — CPU specification of the O(3) tensor product for l <= 2 (groundwork for SURVEY 8f-3,
BASELINE configs[2]: "SEGNN l_max=2").

PARITY UNPINNED for everything that involves l = 2: the reference mount contains no l = 2 code
(`/root/reference/models/segnn/l1_tensor_prod.py:13-14` asserts lmax == 1; e3nn is not installed).  What IS pinned:
restricted to l <= 1 this module reproduces the reference `L1TensorProduct` exactly (same Clebsch-Gordan constants
`L1TP:91-94`, same 'component' x 'element' normalisation `L1TP:122-189`, same weight semantics), which
`tests/test_oracle_lmax2.py` checks against `oracle/l1tp_oracle.py` (itself pinned by golden vectors of the
unmodified reference).  The l = 2 part is fixed by that convention plus SO(3)/O(3) equivariance:

* irreps are real; l = 1 in Cartesian order (x, y, z) (the reference treats `in2[:, 1:4]` as a 3-vector, `L1TP:246-279`);
  l = 2 in the orthonormal basis Q_a of symmetric traceless 3x3 tensors listed in `Q2`;
* the coupling tensor C(l1,l2,l3) is THE invariant tensor of D_l1 x D_l2 x D_l3 (one-dimensional for |l1-l2| <= l3 <=
  l1+l2), computed numerically as the null space of the invariance equations, with unit Frobenius norm (as the
  reference's constants: 1, 1/sqrt3, 1/sqrt6) and the sign of the reference for l <= 1 / of `_sign_convention`
  otherwise;
* fully connected ("uvw") weights per path, output = a_out * sum_paths sum_u W[u,w] sum_ij C[i,j,k] x1[u,i] x2[j],
  a_out = sqrt((2 l_out + 1) / sum_paths mul1 mul2)  ('component' x 'element', `L1TP:124,145,169`).

Only `tests/` may import this file.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from .l1tp_oracle import parse_irreps, irreps_dim

# orthonormal basis of symmetric traceless 3x3 tensors (<Qa, Qb>_F = delta_ab)
_s2, _s6 = 1.0 / math.sqrt(2.0), 1.0 / math.sqrt(6.0)
Q2 = np.zeros((5, 3, 3))
Q2[0, 0, 1] = Q2[0, 1, 0] = _s2                      # xy
Q2[1, 1, 2] = Q2[1, 2, 1] = _s2                      # yz
Q2[2, 0, 0] = Q2[2, 1, 1] = -_s6; Q2[2, 2, 2] = 2 * _s6   # 2zz - xx - yy
Q2[3, 0, 2] = Q2[3, 2, 0] = _s2                      # zx
Q2[4, 0, 0] = _s2; Q2[4, 1, 1] = -_s2                # xx - yy


def wigner_D(l: int, R: np.ndarray) -> np.ndarray:
    """Real representation matrix of the proper rotation R (3x3) on the l-th irrep in the bases above."""
    if l == 0:
        return np.ones((1, 1))
    if l == 1:
        return np.asarray(R, dtype=np.float64)
    if l == 2:
        return np.einsum("aij,ik,bkl,jl->ab", Q2, R, Q2, R)
    raise ValueError("l <= 2")


def _rand_rot(rng) -> np.ndarray:
    q, r = np.linalg.qr(rng.standard_normal((3, 3)))
    q = q * np.sign(np.diag(r))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


_CG_CACHE: Dict[Tuple[int, int, int], np.ndarray] = {}


def cg(l1: int, l2: int, l3: int) -> np.ndarray:
    """Unit-norm invariant tensor C[i, j, k] of (l1 x l2 -> l3); zeros if the triangle rule fails."""
    key = (l1, l2, l3)
    if key in _CG_CACHE:
        return _CG_CACHE[key]
    d1, d2, d3 = 2 * l1 + 1, 2 * l2 + 1, 2 * l3 + 1
    if not (abs(l1 - l2) <= l3 <= l1 + l2):
        c = np.zeros((d1, d2, d3))
    else:
        rng = np.random.default_rng(12345)
        rows = []
        for _ in range(6):
            R = _rand_rot(rng)
            m = np.einsum("ai,bj,ck->abcijk", wigner_D(l1, R), wigner_D(l2, R), wigner_D(l3, R)).reshape(d1 * d2 * d3, -1)
            rows.append(m - np.eye(d1 * d2 * d3))
        _, s, vt = np.linalg.svd(np.concatenate(rows, 0))
        assert s[-1] < 1e-10 and (len(s) == 1 or s[-2] > 1e-6), \
            f"invariant subspace of {key} is not one-dimensional: {s[-3:]}"
        c = vt[-1].reshape(d1, d2, d3)
        c /= np.linalg.norm(c)
        ref = _reference_cg(l1, l2, l3)
        if ref is not None:
            c = c * np.sign((c * ref).sum())
            assert np.abs(c - ref).max() < 1e-10, f"l<=1 coupling {key} differs from the reference constants"
        else:
            c = c * np.sign((c * _sign_convention(l1, l2, l3)).sum())
    _CG_CACHE[key] = c
    return c


def _reference_cg(l1, l2, l3):
    """The reference's hand-written couplings (`L1TP:91-94` with `L1TP:246-295`): delta / sqrt3, epsilon / sqrt6."""
    if max(l1, l2, l3) > 1:
        return None
    c3, c6 = 1.0 / math.sqrt(3.0), 1.0 / math.sqrt(6.0)
    eye = np.eye(3)
    if (l1, l2, l3) == (0, 0, 0):
        return np.ones((1, 1, 1))
    if (l1, l2, l3) == (1, 1, 0):
        return (c3 * eye)[:, :, None]
    if (l1, l2, l3) == (0, 1, 1):
        return (c3 * eye)[None, :, :]
    if (l1, l2, l3) == (1, 0, 1):
        return (c3 * eye)[:, None, :]
    if (l1, l2, l3) == (1, 1, 1):
        eps = np.zeros((3, 3, 3))
        eps[0, 1, 2] = eps[1, 2, 0] = eps[2, 0, 1] = 1.0
        eps[0, 2, 1] = eps[2, 1, 0] = eps[1, 0, 2] = -1.0
        return c6 * eps          # out_k = (in1 x in2)_k / sqrt6, `L1TP:279,293`
    return None


def _sign_convention(l1, l2, l3):
    """The invariant is unique up to sign; the sign chosen for triples involving l = 2: +delta when one factor is a
    scalar, otherwise trace(B1_i B2_j B3_k) with B the orthonormal 3x3 bases (l=1: antisymmetric eps_iab / sqrt2, l=2:
    Q2).  For 1x1->1 this rule gives +epsilon, i.e. it continues the reference's choice."""
    d1, d2, d3 = 2 * l1 + 1, 2 * l2 + 1, 2 * l3 + 1
    if l1 == 0:
        return np.eye(d3)[None, :, :]
    if l2 == 0:
        return np.eye(d3)[:, None, :]
    if l3 == 0:
        return np.eye(d1)[:, :, None]
    eps = np.zeros((3, 3, 3))
    eps[0, 1, 2] = eps[1, 2, 0] = eps[2, 0, 1] = 1.0
    eps[0, 2, 1] = eps[2, 1, 0] = eps[1, 0, 2] = -1.0
    B = {1: eps / math.sqrt(2.0), 2: Q2}
    return np.einsum("iab,jbc,kca->ijk", B[l1], B[l2], B[l3])


def spherical_harmonics(vec: np.ndarray, lmax: int) -> np.ndarray:
    """[E, sum(2l+1)] real SH of the direction of `vec`, 'integral' normalisation (Y_0 = 1/sqrt(4 pi), Y_1 =
    sqrt(3/4pi) n: the values the graph builder writes into edge_attr), l = 2 in the basis Q2."""
    vec = np.asarray(vec, dtype=np.float64)
    r = np.linalg.norm(vec, axis=1, keepdims=True)
    n = np.where(r > 0, vec / np.maximum(r, 1e-300), 0.0)
    out = [np.full((len(vec), 1), 1.0 / math.sqrt(4 * math.pi))]
    if lmax >= 1:
        out.append(math.sqrt(3.0 / (4 * math.pi)) * n)
    if lmax >= 2:
        out.append(math.sqrt(5.0 / (4 * math.pi)) * math.sqrt(1.5) * np.einsum("aij,ei,ej->ea", Q2, n, n))
    return np.concatenate(out, 1)


def sh_irreps(lmax: int) -> List[Tuple[int, int, int]]:
    return [(1, l, (-1) ** l) for l in range(lmax + 1)]


def paths(in1, in2, out) -> List[Tuple[int, int, int]]:
    """(i_in1, i_in2, i_out) in the reference's enumeration order (io, ii2, ii1) (`L1TP:122-151`), proper selection
    rules (triangle + parity).  NOTE: the reference's quirk Q1 (parity not checked for l = 0 outputs) is NOT
    reproduced here; it is invisible for SH-type inputs."""
    ps = []
    for io, (_, lo, po) in enumerate(out):
        for i2, (_, l2, p2) in enumerate(in2):
            for i1, (_, l1, p1) in enumerate(in1):
                if abs(l1 - l2) <= lo <= l1 + l2 and po == p1 * p2:
                    ps.append((i1, i2, io))
    return ps


def weight_shapes(in1, in2, out) -> List[Tuple[int, int]]:
    return [(in1[i1][0] * in2[i2][0], out[io][0]) for i1, i2, io in paths(in1, in2, out)]


def norm_factors(in1, in2, out) -> List[float]:
    a = []
    ps = paths(in1, in2, out)
    for io, (_, lo, _) in enumerate(out):
        x = sum(in1[i1][0] * in2[i2][0] for i1, i2, o in ps if o == io)
        a.append(math.sqrt((2 * lo + 1) / x) if x > 0 else 0.0)
    return a


def _offsets(irreps):
    off, o = [], 0
    for mul, l, _ in irreps:
        off.append(o)
        o += mul * (2 * l + 1)
    return off


def forward(x1: torch.Tensor, x2: torch.Tensor, weights: Sequence[torch.Tensor], in1, in2, out) -> torch.Tensor:
    """x1 [E, dim(in1)], x2 [E, dim(in2)] (mul 1 per in2 irrep), weights per path [mul1, mul_out] -> [E, dim(out)].
    Differentiable (torch fp64): the backward of the specification is autograd's."""
    in1, in2, out = parse_irreps(in1) if isinstance(in1, str) else in1, parse_irreps(in2) if isinstance(in2, str) else in2, \
        parse_irreps(out) if isinstance(out, str) else out
    o1, o2, oo = _offsets(in1), _offsets(in2), _offsets(out)
    a = norm_factors(in1, in2, out)
    E = x1.shape[0]
    res = [torch.zeros((E, mo, 2 * lo + 1), dtype=x1.dtype) for mo, lo, _ in out]
    for (i1, i2, io), W in zip(paths(in1, in2, out), weights):
        m1, l1, _ = in1[i1]
        m2, l2, _ = in2[i2]
        assert m2 == 1, "in2 is a spherical-harmonics type input (mul 1)"
        C = torch.from_numpy(cg(l1, l2, out[io][1])).to(x1.dtype)
        a1 = x1[:, o1[i1]:o1[i1] + m1 * (2 * l1 + 1)].reshape(E, m1, 2 * l1 + 1)
        a2 = x2[:, o2[i2]:o2[i2] + (2 * l2 + 1)]
        f = torch.einsum("eui,ej,ijk->euk", a1, a2, C)
        res[io] = res[io] + a[io] * torch.einsum("euk,uw->ewk", f, W)
    return torch.cat([r.reshape(E, -1) for r in res], 1)


def weights_from_l1tp(in1, out, w: Dict[str, np.ndarray]) -> List[np.ndarray]:
    """Per-path weights equivalent to the reference's stacked `weights_l0e/l0o/l1e/l1o` (`L1TP:81-88`: rows in the
    `cat` order of forward, columns = output channels of the species) for in2 = SH(1)."""
    in2 = sh_irreps(1)
    sp = lambda l, p: f"{l}{'e' if p == 1 else 'o'}"
    # row blocks of each stacked matrix: species order of the `cat`s in L1TP:244-295
    cat_order = {"0e": ["0e", "1o"], "0o": ["0o", "1e"], "1e": ["0o", "1e", "1o"], "1o": ["0e", "1o", "1e"]}
    mul_of = {s: sum(m for m, l, p in in1 if sp(l, p) == s) for s in ("0e", "0o", "1e", "1o")}
    ws = []
    for i1, i2, io in paths(in1, in2, out):
        m1, l1, p1 = in1[i1]
        mo, lo, po = out[io]
        so, s1 = sp(lo, po), sp(l1, p1)
        W = w[f"weights_l{so}"]
        r0 = 0
        for s in cat_order[so]:
            if s == s1:
                break
            r0 += mul_of[s]
        r0 += sum(m for m, l, p in in1[:i1] if sp(l, p) == s1)           # earlier irreps of the same species
        c0 = sum(m for m, l, p in out[:io] if sp(l, p) == so)
        ws.append(np.asarray(W[r0:r0 + m1, c0:c0 + mo]))
    return ws
