"""TEST INFRASTRUCTURE ONLY — CPU oracle for the l<=1 Clebsch-Gordan tensor product.

A numpy restatement of the algorithm in the reference file
``/root/reference/models/segnn/l1_tensor_prod.py`` (cited below as ``L1TP:<line>``).
It is the *checker* for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.
The product package never does.

Pinning: ``tests/golden/l1tp_*.npz`` were produced by running the unmodified
reference file in the build container (``tests/golden/make_l1tp_golden.py``);
``tests/test_oracle_l1tp.py`` checks this restatement against every one of them
(fp64 to ~1e-13, fp32 to ~1e-6), plus the analytic norm values from SURVEY §8c.

Layout conventions (L1TP:24-36, 247, 276): features are e3nn-flat, irreps in
declaration order, an ``mul x 1p`` block is ``[mul, 3]`` row-major.  Same-species
blocks of interleaved irreps are concatenated in order (boolean-mask semantics).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

C3 = 1.0 / math.sqrt(3.0)  # cg110 = cg011, L1TP:92-93
C6 = 1.0 / math.sqrt(6.0)  # cg111, L1TP:94

SPECIES = ("0e", "0o", "1e", "1o")
SH1 = ((1, 0, 1), (1, 1, -1))  # Irreps.spherical_harmonics(1) = 1x0e+1x1o, L1TP:17


def parse_irreps(spec) -> List[Tuple[int, int, int]]:
    """'4x0e+2x1o' -> [(mul, l, p)]; also accepts iterables of objects with .mul/.ir."""
    if not isinstance(spec, str):
        return [(int(m.mul), int(m.ir.l), int(m.ir.p)) for m in spec]
    out = []
    for tok in [t.strip() for t in spec.split("+") if t.strip()]:
        mul, _, ir = tok.rpartition("x")
        out.append((int(mul) if mul else 1, int(ir[:-1]), 1 if ir[-1] == "e" else -1))
    return out


def irreps_dim(irreps) -> int:
    return sum(mul * (2 * l + 1) for mul, l, _ in irreps)


def species_columns(irreps) -> Dict[str, np.ndarray]:
    """Flat column indices per species, in declaration order (L1TP:24-36, 53-65).

    For l=0 species: one column per channel.  For l=1 species: the column of the
    x component of every vector (the vector is columns c, c+1, c+2)."""
    cols = {s: [] for s in SPECIES}
    i = 0
    for mul, l, p in irreps:
        key = f"{l}{'e' if p == 1 else 'o'}"
        if l == 0:
            cols[key].extend(range(i, i + mul))
        elif l == 1:
            cols[key].extend(range(i, i + 3 * mul, 3))
        else:
            raise AssertionError("lmax must be 1 (L1TP:13-14)")
        i += mul * (2 * l + 1)
    return {k: np.asarray(v, dtype=np.int64) for k, v in cols.items()}


def weight_shapes(in1, out) -> Dict[str, Tuple[int, int]]:
    """Row order = the `cat` order in forward (L1TP:81-88, 244-295)."""
    ci, co = species_columns(in1), species_columns(out)
    n = {k: len(v) for k, v in ci.items()}
    m = {k: len(v) for k, v in co.items()}
    shapes = {}
    if n["0e"] + n["1o"] > 0 and m["0e"] > 0:
        shapes["weights_l0e"] = (n["0e"] + n["1o"], m["0e"])
    if n["0o"] + n["1e"] > 0 and m["0o"] > 0:
        shapes["weights_l0o"] = (n["0o"] + n["1e"], m["0o"])
    if n["0o"] + n["1e"] + n["1o"] > 0 and m["1e"] > 0:
        shapes["weights_l1e"] = (n["0o"] + n["1e"] + n["1o"], m["1e"])
    if n["0e"] + n["1o"] + n["1e"] > 0 and m["1o"] > 0:
        shapes["weights_l1o"] = (n["0e"] + n["1o"] + n["1e"], m["1o"])
    return shapes


def norm_factors(in1, out, irrep_normalization="component", path_normalization="element",
                 in1_var: Optional[Sequence[float]] = None,
                 in2_var: Optional[Sequence[float]] = None,
                 out_var: Optional[Sequence[float]] = None):
    """Per-output-irrep factor ``a`` and weight-init half-width ``wi`` (L1TP:115-193).

    Reproduces quirk Q1 (L1TP:137-138): `A or (B and C)` precedence, i.e. for l=0
    outputs every (in1.l == in2.l) pair counts regardless of parity.
    Returns (a_list, wi_list, instructions) with instructions as plain tuples
    (i_in1, i_in2, i_out, 'uvw', True, a, (mul1, mul2, mul_out))."""
    iri2 = SH1
    in1_var = [1.0] * len(in1) if in1_var is None else [float(v) for v in in1_var]
    in2_var = [1.0] * len(iri2) if in2_var is None else [float(v) for v in in2_var]
    out_var = [1.0] * len(out) if out_var is None else [float(v) for v in out_var]
    assert len(in1_var) == len(in1) and len(in2_var) == len(iri2) and len(out_var) == len(out)
    if irrep_normalization not in ("component", "none") or path_normalization not in ("element", "none"):
        raise AssertionError("Not all norms are implemented yet.")  # L1TP:117-118
    if irrep_normalization == "none" and path_normalization == "none":
        raise AttributeError("is_comp_norm undefined (reference quirk Q3, L1TP:116)")
    a_list, wi_list, instr = [], [], []
    for io, (mo, lo, po) in enumerate(out):
        alpha = (2 * lo + 1) * out_var[io] if irrep_normalization == "component" else 1.0
        x = 0.0
        paths = []
        for ii2, (m2, l2, p2) in enumerate(iri2):
            for ii1, (m1, l1, p1) in enumerate(in1):
                hit = (lo == 0 and l2 == l1) or (lo == 1 and (l2 | l1) and po == p2 * p1)
                if hit:
                    x += in1_var[ii1] * in2_var[ii2] * m1 * m2
                    paths.append((ii1, ii2, io, "uvw", True, None, (m1, m2, mo)))
        if path_normalization == "none":
            a = math.sqrt(alpha)
            wi = 1.0 / math.sqrt(x)
        else:
            a = math.sqrt(alpha / x) if x > 0 else math.sqrt(alpha)
            wi = 1.0
        a_list.append(a)
        wi_list.append(wi)
        instr.extend(p[:5] + (a,) + p[6:] for p in paths)
    return a_list, wi_list, instr


def norm_buffers(out, a_list) -> Dict[str, np.ndarray]:
    """norm_l0e/l0o/l1e/l1o buffers; l=1 buffers have 3 entries per channel (L1TP:159-189)."""
    bufs = {s: [] for s in SPECIES}
    for (mo, lo, po), a in zip(out, a_list):
        key = f"{lo}{'e' if po == 1 else 'o'}"
        bufs[key].extend([a] * (mo * (2 * lo + 1)))
    return {f"norm_l{k}": np.asarray(v, dtype=np.float64) for k, v in bufs.items()}


def _vec(x, cols):
    """[E, n, 3] view of the l=1 channels whose x-columns are `cols`."""
    if len(cols) == 0:
        return np.zeros((x.shape[0], 0, 3), dtype=x.dtype)
    return np.stack([x[:, cols], x[:, cols + 1], x[:, cols + 2]], axis=-1)


def features(in1: np.ndarray, in2: np.ndarray, ci) -> Dict[str, np.ndarray]:
    """The four `cat`-ed feature blocks of the forward (L1TP:244-250, 260-266, 274-281, 288-295).

    f0e,f0o: [E,K];  f1e,f1o: [E,K,3]."""
    y0 = in2[:, 0:1]
    y1 = in2[:, None, 1:4]
    s0e, s0o = in1[:, ci["0e"]], in1[:, ci["0o"]]
    v1e, v1o = _vec(in1, ci["1e"]), _vec(in1, ci["1o"])
    f = {}
    f["0e"] = np.concatenate([s0e * y0, C3 * (v1o * y1).sum(-1)], axis=1)
    f["0o"] = np.concatenate([s0o * y0, C3 * (v1e * y1).sum(-1)], axis=1)
    f["1e"] = np.concatenate([C3 * s0o[:, :, None] * y1, C3 * v1e * y0[:, :, None],
                              C6 * np.cross(v1o, np.broadcast_to(y1, v1o.shape))], axis=1)
    f["1o"] = np.concatenate([C3 * s0e[:, :, None] * y1, C3 * v1o * y0[:, :, None],
                              C6 * np.cross(v1e, np.broadcast_to(y1, v1e.shape))], axis=1)
    return f


def forward(in1, in2, weights: Dict[str, np.ndarray], norms: Dict[str, np.ndarray], in1_irreps, out_irreps):
    """out[E, Dout] (L1TP:234-299).  Columns of species with zero width / no weights stay 0."""
    in1_irreps, out_irreps = parse_irreps(in1_irreps), parse_irreps(out_irreps)
    ci, co = species_columns(in1_irreps), species_columns(out_irreps)
    f = features(in1, in2, ci)
    out = np.zeros((in1.shape[0], irreps_dim(out_irreps)), dtype=in1.dtype)
    for sp in ("0e", "0o"):
        w = weights.get(f"weights_l{sp}")
        if len(co[sp]) and w is not None:
            out[:, co[sp]] = (f[sp] @ w) * norms[f"norm_l{sp}"].astype(in1.dtype)
    for sp in ("1e", "1o"):
        w = weights.get(f"weights_l{sp}")
        if len(co[sp]) and w is not None:
            r = np.einsum("ekc,km->emc", f[sp], w)  # tensordot + transpose, L1TP:281,295
            r = r * norms[f"norm_l{sp}"].astype(in1.dtype).reshape(1, -1, 3)
            for c in range(3):
                out[:, co[sp] + c] = r[:, :, c]
    return out


def backward(in1, in2, gout, weights, norms, in1_irreps, out_irreps):
    """Closed-form VJP of `forward`: returns (g_in1, g_in2, {g_weights})."""
    in1_irreps, out_irreps = parse_irreps(in1_irreps), parse_irreps(out_irreps)
    ci, co = species_columns(in1_irreps), species_columns(out_irreps)
    n = {k: len(v) for k, v in ci.items()}
    f = features(in1, in2, ci)
    E = in1.shape[0]
    y0 = in2[:, 0:1]
    y1 = in2[:, None, 1:4]
    s = {"0e": in1[:, ci["0e"]], "0o": in1[:, ci["0o"]]}
    v = {"1e": _vec(in1, ci["1e"]), "1o": _vec(in1, ci["1o"])}
    g_in1 = np.zeros_like(in1)
    g_y0 = np.zeros((E, 1), dtype=in1.dtype)
    g_y1 = np.zeros((E, 3), dtype=in1.dtype)
    gs = {k: np.zeros_like(x) for k, x in s.items()}
    gv = {k: np.zeros_like(x) for k, x in v.items()}
    gw = {}
    # l=0 outputs: f = [s*y0, C3*<v,y1>]
    for sp, ssp, vsp in (("0e", "0e", "1o"), ("0o", "0o", "1e")):
        w = weights.get(f"weights_l{sp}")
        if not len(co[sp]) or w is None:
            continue
        g = gout[:, co[sp]] * norms[f"norm_l{sp}"].astype(in1.dtype)
        gw[f"weights_l{sp}"] = f[sp].T @ g
        gf = g @ w.T
        ns = n[ssp]
        gs[ssp] += gf[:, :ns] * y0
        g_y0 += (gf[:, :ns] * s[ssp]).sum(1, keepdims=True)
        gd = gf[:, ns:]
        gv[vsp] += C3 * gd[:, :, None] * y1
        g_y1 += C3 * (gd[:, :, None] * v[vsp]).sum(1)
    # l=1 outputs: f = [C3*s (x) y1, C3*v*y0, C6*(v' x y1)]
    for sp, ssp, vsp, xsp in (("1e", "0o", "1e", "1o"), ("1o", "0e", "1o", "1e")):
        w = weights.get(f"weights_l{sp}")
        if not len(co[sp]) or w is None:
            continue
        g = np.stack([gout[:, co[sp] + c] for c in range(3)], axis=-1)  # [E,m,3]
        g = g * norms[f"norm_l{sp}"].astype(in1.dtype).reshape(1, -1, 3)
        gw[f"weights_l{sp}"] = np.einsum("ekc,emc->km", f[sp], g)
        gf = np.einsum("emc,km->ekc", g, w)
        ns, nv = n[ssp], n[vsp]
        gfs, gfv, gfx = gf[:, :ns], gf[:, ns:ns + nv], gf[:, ns + nv:]
        gs[ssp] += C3 * (gfs * y1).sum(-1)
        g_y1 += C3 * (gfs * s[ssp][:, :, None]).sum(1)
        gv[vsp] += C3 * gfv * y0[:, :, None]
        g_y0 += C3 * (gfv * v[vsp]).sum((1, 2))[:, None]
        # d/dv' of gfx.(v' x y1) = y1 x gfx ;  d/dy1 = gfx x v'
        y1b = np.broadcast_to(y1, gfx.shape)
        gv[xsp] += C6 * np.cross(y1b, gfx)
        g_y1 += C6 * np.cross(gfx, v[xsp]).sum(1)
    for k in ("0e", "0o"):
        g_in1[:, ci[k]] = gs[k]
    for k in ("1e", "1o"):
        for c in range(3):
            g_in1[:, ci[k] + c] = gv[k][:, :, c]
    g_in2 = np.concatenate([g_y0, g_y1], axis=1)
    return g_in1, g_in2, gw
