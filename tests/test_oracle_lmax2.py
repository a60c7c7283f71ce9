"""The l <= 2 specification (`oracle/lmax2_oracle.py`, groundwork for the l_max = 2 path): restricted to l <= 1 it must be
the reference L1TensorProduct (through `oracle/l1tp_oracle.py`, which the golden vectors pin); for l = 2 it must be
O(3)-equivariant with unit-norm couplings."""
import numpy as np
import pytest
import torch

from oracle import l1tp_oracle as l1
from oracle import lmax2_oracle as l2


def _rand_weights_l1(in1, out, rng):
    return {k: rng.standard_normal(s) for k, s in l1.weight_shapes(in1, out).items()}


@pytest.mark.parametrize("in1,out", [
    ("8x0e+4x1o", "6x0e+5x1o"),
    ("3x0e+2x1o+2x0e+1x1o", "4x0e+2x1o+3x0e"),      # interleaved species: row offsets inside the stacked weights
    ("5x0e", "2x0e+3x1o"),
    ("4x1o", "3x0e+2x1o"),
    ("2x0e+3x1o+2x1e+1x0o", "2x1o+2x1e"),            # l = 1 outputs only: quirk Q1 (l = 0 outputs) not involved
])
def test_restriction_to_l1_is_the_reference(in1, out):
    rng = np.random.default_rng(0)
    ir1, iro = l1.parse_irreps(in1), l1.parse_irreps(out)
    E = 37
    x1 = rng.standard_normal((E, l1.irreps_dim(ir1)))
    x2 = l2.spherical_harmonics(rng.standard_normal((E, 3)), 1)
    w = _rand_weights_l1(ir1, iro, rng)
    a, _, _ = l1.norm_factors(ir1, iro)
    want = l1.forward(x1, x2, w, l1.norm_buffers(iro, a), in1, out)
    ws = [torch.from_numpy(np.ascontiguousarray(m)) for m in l2.weights_from_l1tp(ir1, iro, w)]
    got = l2.forward(torch.from_numpy(x1), torch.from_numpy(x2), ws, ir1, l2.sh_irreps(1), iro).numpy()
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(l2.norm_factors(ir1, l2.sh_irreps(1), iro), a, rtol=1e-14)


def test_couplings_unit_norm_and_selection():
    for a in range(3):
        for b in range(3):
            for c in range(3):
                C = l2.cg(a, b, c)
                assert C.shape == (2 * a + 1, 2 * b + 1, 2 * c + 1)
                if abs(a - b) <= c <= a + b:
                    assert abs(np.linalg.norm(C) - 1.0) < 1e-12
                else:
                    assert not C.any()
    # the l <= 1 entries are the reference's constants
    assert abs(l2.cg(1, 1, 0)[0, 0, 0] - 1 / np.sqrt(3)) < 1e-12
    assert abs(l2.cg(1, 1, 1)[0, 1, 2] - 1 / np.sqrt(6)) < 1e-12


def test_sh_norms_and_rotation():
    rng = np.random.default_rng(1)
    v = rng.standard_normal((50, 3))
    y = l2.spherical_harmonics(v, 2)
    # 'integral' normalisation: sum_m Y_lm^2 = (2l+1)/(4 pi)
    np.testing.assert_allclose((y[:, 0:1] ** 2).sum(1), 1 / (4 * np.pi), rtol=1e-12)
    np.testing.assert_allclose((y[:, 1:4] ** 2).sum(1), 3 / (4 * np.pi), rtol=1e-12)
    np.testing.assert_allclose((y[:, 4:9] ** 2).sum(1), 5 / (4 * np.pi), rtol=1e-12)
    R = l2._rand_rot(rng)
    yr = l2.spherical_harmonics(v @ R.T, 2)
    np.testing.assert_allclose(yr[:, 1:4], y[:, 1:4] @ l2.wigner_D(1, R).T, atol=1e-12)
    np.testing.assert_allclose(yr[:, 4:9], y[:, 4:9] @ l2.wigner_D(2, R).T, atol=1e-12)
    # a zero vector (self edge) gives Y_0 only, as the graph builder's l = 1 attributes do
    z = l2.spherical_harmonics(np.zeros((1, 3)), 2)
    assert z[0, 0] > 0 and not z[0, 1:].any()


def _block_D(irreps, R, inversion):
    mats = []
    for mul, l, p in irreps:
        D = l2.wigner_D(l, R) * (p if inversion else 1)
        mats.extend([D] * mul)
    n = sum(m.shape[0] for m in mats)
    out = np.zeros((n, n))
    o = 0
    for m in mats:
        out[o:o + len(m), o:o + len(m)] = m
        o += len(m)
    return out


@pytest.mark.parametrize("inversion", [False, True])
def test_lmax2_equivariance_and_grad(inversion):
    rng = np.random.default_rng(2)
    in1 = [(4, 0, 1), (3, 1, -1), (2, 2, 1), (1, 1, 1), (1, 2, -1), (1, 0, -1)]
    out = [(3, 0, 1), (2, 1, -1), (2, 2, 1), (1, 1, 1), (1, 2, -1), (1, 0, -1)]
    in2 = l2.sh_irreps(2)
    E = 11
    d1 = sum(m * (2 * l + 1) for m, l, _ in in1)
    x1 = rng.standard_normal((E, d1))
    vec = rng.standard_normal((E, 3))
    ws = [torch.from_numpy(rng.standard_normal(s)).requires_grad_() for s in l2.weight_shapes(in1, in2, out)]
    assert len(ws) == len(l2.paths(in1, in2, out)) > 20
    R = l2._rand_rot(rng)
    sgn = -1.0 if inversion else 1.0
    y = torch.from_numpy(l2.spherical_harmonics(vec, 2))
    yr = torch.from_numpy(l2.spherical_harmonics(sgn * vec @ R.T, 2))
    x1t = torch.from_numpy(x1).requires_grad_()
    o = l2.forward(x1t, y, ws, in1, in2, out)
    o_rot = l2.forward(torch.from_numpy(x1 @ _block_D(in1, R, inversion).T), yr, ws, in1, in2, out)
    np.testing.assert_allclose(o_rot.detach().numpy(), o.detach().numpy() @ _block_D(out, R, inversion).T, atol=1e-11)
    # autograd backward of the specification against central differences on one weight and one input entry
    g = torch.from_numpy(rng.standard_normal(o.shape))
    (o * g).sum().backward()
    eps = 1e-6
    with torch.no_grad():
        w0 = ws[3]
        w0[0, 0] += eps
        fp = (l2.forward(x1t, y, ws, in1, in2, out) * g).sum()
        w0[0, 0] -= 2 * eps
        fm = (l2.forward(x1t, y, ws, in1, in2, out) * g).sum()
        w0[0, 0] += eps
    assert abs((fp - fm).item() / (2 * eps) - ws[3].grad[0, 0].item()) < 1e-6
    assert x1t.grad is not None and torch.isfinite(x1t.grad).all()


def test_component_normalisation_variance():
    """'component' x 'element': unit-variance inputs and weights, component-normalised SH -> O(1) outputs (the property
    the reference's normalisation is built for, `L1TP:124-151`)."""
    rng = np.random.default_rng(3)
    in1 = [(16, 0, 1), (16, 1, -1), (16, 2, 1)]
    out = [(8, 0, 1), (8, 1, -1), (8, 2, 1)]
    in2 = l2.sh_irreps(2)
    E = 4000
    x1 = torch.from_numpy(rng.standard_normal((E, 16 * 9)))
    y = l2.spherical_harmonics(rng.standard_normal((E, 3)), 2) * np.sqrt(4 * np.pi)      # component normalisation
    ws = [torch.from_numpy(rng.standard_normal(s)) for s in l2.weight_shapes(in1, in2, out)]
    o = l2.forward(x1, torch.from_numpy(y), ws, in1, in2, out).numpy()
    for lo, (a, b) in zip((0, 1, 2), ((0, 8), (8, 32), (32, 72))):
        v = o[:, a:b].var()
        assert 0.5 < v < 2.0, (lo, v)


def test_frozen_convention():
    """`tests/golden/o3tp_*.npz` freeze this repo's l = 2 convention (bases, coupling signs, normalisation, path order):
    the oracle must keep reproducing them.  (Not reference outputs: there is no l = 2 reference.)"""
    import glob
    import json
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    z = np.load(os.path.join(here, "o3tp_couplings.npz"))
    assert len(z.files) == 15
    for k in z.files:
        a, b, c = (int(ch) for ch in k[3:])
        np.testing.assert_allclose(l2.cg(a, b, c), z[k], atol=1e-13)
    files = sorted(f for f in glob.glob(os.path.join(here, "o3tp_*.npz")) if not f.endswith("couplings.npz"))
    assert len(files) == 3
    for f in files:
        r = np.load(f)
        meta = json.loads(bytes(r["meta"]).decode())
        in1 = [tuple(t) for t in meta["in1"]]
        out = [tuple(t) for t in meta["out"]]
        in2 = l2.sh_irreps(meta["lmax"])
        assert [list(p) for p in l2.paths(in1, in2, out)] == meta["paths"]
        np.testing.assert_allclose(l2.norm_factors(in1, in2, out), meta["norm"], rtol=1e-14)
        ws, o = [], 0
        for a, b in l2.weight_shapes(in1, in2, out):
            ws.append(torch.from_numpy(r["w"][o:o + a * b].astype(np.float64).reshape(a, b)))
            o += a * b
        got = l2.forward(torch.from_numpy(r["x1"].astype(np.float64)), torch.from_numpy(r["y"].astype(np.float64)), ws,
                         in1, in2, out).numpy()
        np.testing.assert_allclose(got, r["out_f64"], rtol=1e-12, atol=1e-12)
