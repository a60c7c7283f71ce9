"""GPU parity of the l <= 2 tensor product (`se3_o3tp_*` through the C ABI / `O3TensorProduct`) against
(1) the fp64 specification `oracle/lmax2_oracle.py` (forward and autograd backward) on seeded inputs spanning many
    tiles and CTAs with ragged tails,
(2) the l <= 1 product path (`L1TensorProduct`, pinned by the reference's golden vectors) on SH(1)-type irreps, and
(3) at 300k rows, the size-independent properties: O(3) equivariance and linearity of the weight gradient.
Tolerance 1e-5 relative to the largest reference magnitude, fp32."""
import numpy as np
import pytest
import torch

from oracle import l1tp_oracle as O1
from oracle import lmax2_oracle as O2

pytestmark = pytest.mark.gpu

RTOL = 1e-5

CASES = {
    "sh1": ([(8, 0, 1), (4, 1, -1)], 1, [(6, 0, 1), (5, 1, -1)]),
    "balanced2": ([(23, 0, 1), (7, 1, -1), (4, 2, 1)], 2, [(23, 0, 1), (7, 1, -1), (4, 2, 1)]),
    "message2": ([(23, 0, 1), (7, 1, -1), (4, 2, 1), (23, 0, 1), (7, 1, -1), (4, 2, 1), (2, 0, 1)], 2,
                 [(34, 0, 1), (7, 1, -1), (4, 2, 1)]),
    "mixed_parity": ([(3, 0, 1), (2, 1, -1), (2, 2, 1), (1, 1, 1), (1, 2, -1), (2, 0, -1)], 2,
                     [(3, 0, 1), (2, 1, -1), (1, 2, 1), (2, 1, 1), (1, 2, -1), (5, 0, -1)]),
    "dead_output": ([(4, 0, 1)], 1, [(3, 0, 1), (2, 1, 1), (2, 1, -1)]),
    "scalar_attr": ([(5, 0, 1), (3, 2, 1)], 0, [(4, 0, 1), (2, 2, 1)]),
    "wide": ([(64, 0, 1), (32, 1, -1), (16, 2, 1)], 2, [(64, 0, 1), (32, 1, -1), (16, 2, 1)]),
}


def _spec(ir):
    return "+".join(f"{m}x{l}{'e' if p == 1 else 'o'}" for m, l, p in ir)


def _module(in1, lmax, out, seed=0):
    from se3gnn_b200.irreps import Irreps
    from se3gnn_b200.o3tp import O3TensorProduct
    torch.manual_seed(seed)
    return O3TensorProduct(Irreps(_spec(in1)), Irreps(_spec(out)), Irreps.spherical_harmonics(lmax)).cuda()


def _close(a, ref, what, rtol=RTOL):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert a.shape == ref.shape, (what, a.shape, ref.shape)
    err = np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-30)
    assert err <= rtol, f"{what}: rel err {err:.3e} > {rtol}"


def _oracle(in1, in2, out, x1, y, w, g):
    ws, o = [], 0
    for shp in O2.weight_shapes(in1, in2, out):
        n = shp[0] * shp[1]
        ws.append(torch.from_numpy(w[o:o + n].astype(np.float64).reshape(shp)).requires_grad_())
        o += n
    x1t = torch.from_numpy(x1.astype(np.float64)).requires_grad_()
    yt = torch.from_numpy(y.astype(np.float64)).requires_grad_()
    res = O2.forward(x1t, yt, ws, in1, in2, out)
    (res * torch.from_numpy(g.astype(np.float64))).sum().backward()
    return res.detach().numpy(), x1t.grad.numpy(), yt.grad.numpy(), np.concatenate([t.grad.numpy().reshape(-1) for t in ws])


def test_library_couplings_are_the_oracles():
    from se3gnn_b200 import capi
    for a in range(3):
        for b in range(3):
            for c in range(3):
                want = O2.cg(a, b, c)
                if not want.any():
                    with pytest.raises(capi.Se3Error):
                        capi.o3tp_coupling(a, b, c)
                    continue
                np.testing.assert_allclose(np.asarray(capi.o3tp_coupling(a, b, c)), want, atol=1e-12)


@pytest.mark.parametrize("name,rows", [("sh1", 4099), ("balanced2", 10007), ("message2", 6001), ("mixed_parity", 3000),
                                       ("dead_output", 515), ("scalar_attr", 1000), ("wide", 2049), ("balanced2", 1),
                                       ("message2", 31)])
def test_matches_oracle(name, rows):
    from se3gnn_b200 import capi
    in1, lmax, out = CASES[name]
    in2 = O2.sh_irreps(lmax)
    tp = _module(in1, lmax, out)
    assert [(i.i_in1, i.i_in2, i.i_out) for i in tp.instructions] == O2.paths(in1, in2, out)
    rng = np.random.default_rng(rows)
    x1 = rng.standard_normal((rows, tp.in1_dim)).astype(np.float32)
    y = rng.standard_normal((rows, tp.in2_dim)).astype(np.float32)
    g = rng.standard_normal((rows, tp.iro.dim)).astype(np.float32)
    w = tp.weight.detach().cpu().numpy()
    want_o, want_gx, want_gy, want_gw = _oracle(in1, in2, out, x1, y, w, g)
    n0 = capi.launch_count()
    xt = torch.from_numpy(x1).cuda().requires_grad_()
    yt = torch.from_numpy(y).cuda().requires_grad_()
    res = tp(xt, yt)
    assert res.is_contiguous() and res.dtype == torch.float32 and res.shape == want_o.shape
    res.backward(torch.from_numpy(g).cuda())
    torch.cuda.synchronize()
    # one forward kernel; backward = input-gradient + weight-gradient kernels, or the fused fallback kernel; with the
    # weight gradient on the tensor cores (d_in1 <= 64): input-gradient kernel + tcgen05 kernel + its reduction, and the
    # SIMT kernel once more for the rows beyond the last whole 32-row tile
    tc0 = capi.tc_launch_count()
    if tp._plan.tc_weight_grad and rows >= 32:
        assert capi.launch_count() - n0 == 4 + (1 if rows % 32 else 0)
    else:
        assert capi.launch_count() - n0 == (3 if tp._plan.split_backward else 2)
    assert tp._plan.split_backward == (name not in ("mixed_parity", "wide"))
    _close(res, want_o, "out")
    _close(xt.grad, want_gx, "grad in1")
    _close(yt.grad, want_gy, "grad in2")
    _close(tp.weight.grad, want_gw, "grad weight")
    # without a gradient for in2 the kernel skips it; the other results are unchanged
    tp.weight.grad = None
    xt2 = torch.from_numpy(x1).cuda().requires_grad_()
    tp(xt2, torch.from_numpy(y).cuda()).backward(torch.from_numpy(g).cuda())
    _close(xt2.grad, want_gx, "grad in1 (no gin2)")
    _close(tp.weight.grad, want_gw, "grad weight (no gin2)")


def test_zero_rows_and_errors():
    from se3gnn_b200 import capi
    tp = _module(*CASES["balanced2"])
    x = torch.zeros((0, tp.in1_dim), device="cuda", requires_grad=True)
    y = torch.zeros((0, tp.in2_dim), device="cuda")
    o = tp(x, y)
    assert o.shape == (0, tp.iro.dim)
    o.sum().backward()
    assert tp.weight.grad is not None and not tp.weight.grad.any()
    with pytest.raises(Exception):
        tp(torch.zeros((4, tp.in1_dim + 1), device="cuda"), torch.zeros((4, tp.in2_dim), device="cuda"))
    with pytest.raises(capi.Se3Error):
        tp(torch.zeros((4, tp.in1_dim)), torch.zeros((4, tp.in2_dim)))
    with pytest.raises(capi.Se3Error):
        capi.O3tpPlan([(4, 0, 1)], [(1, -1)], [(4, 2, 1)])          # no path
    with pytest.raises(capi.Se3Error):
        capi.O3tpPlan([(4, 3, -1)], [(0, 1)], [(4, 3, -1)])         # l = 3


@pytest.mark.parametrize("in1,out", [("34x0e+10x1o", "34x0e+10x1o"), ("3x0e+2x1o+2x0e+1x1o", "4x0e+2x1o+3x0e")])
def test_l1_restriction_equals_l1_product_path(in1, out):
    """On SH(1)-type irreps the general kernel and the reference-pinned L1TensorProduct kernels agree."""
    from se3gnn_b200.irreps import Irreps
    from models.segnn.l1_tensor_prod import L1TensorProduct
    from se3gnn_b200.o3tp import O3TensorProduct
    torch.manual_seed(0)
    ref = L1TensorProduct(Irreps(in1), Irreps(out)).cuda()
    tp = O3TensorProduct(Irreps(in1), Irreps(out), Irreps.spherical_harmonics(1)).cuda()
    sd = {k: v.detach().cpu().double().numpy() for k, v in ref.state_dict().items() if k.startswith("weights_")}
    ws = O2.weights_from_l1tp(O1.parse_irreps(in1), O1.parse_irreps(out), sd)
    with torch.no_grad():
        tp.weight.copy_(torch.from_numpy(np.concatenate([np.ascontiguousarray(m).reshape(-1) for m in ws])).float())
    rows = 7777
    x = torch.randn(rows, tp.in1_dim, device="cuda")
    y = torch.from_numpy(O2.spherical_harmonics(np.random.default_rng(0).standard_normal((rows, 3)), 1)).float().cuda()
    _close(tp(x, y), ref(x, y).detach().cpu().double().numpy(), "l<=1 restriction", rtol=2e-5)


def _block_D(irreps, R, inversion):
    mats = []
    for mul, l, p in irreps:
        mats.extend([O2.wigner_D(l, R) * (p if inversion else 1)] * mul)
    n = sum(m.shape[0] for m in mats)
    out = np.zeros((n, n))
    o = 0
    for m in mats:
        out[o:o + len(m), o:o + len(m)] = m
        o += len(m)
    return out


def test_large_equivariance_and_gradient_linearity():
    in1, lmax, out = CASES["message2"]
    tp = _module(in1, lmax, out)
    rows = 300_000
    rng = np.random.default_rng(5)
    R = O2._rand_rot(rng)
    x = torch.randn(rows, tp.in1_dim, device="cuda")
    vec = rng.standard_normal((rows, 3))
    y = torch.from_numpy(O2.spherical_harmonics(vec, 2)).float().cuda()
    yr = torch.from_numpy(O2.spherical_harmonics(-vec @ R.T, 2)).float().cuda()       # rotation + inversion
    D1 = torch.from_numpy(_block_D(in1, R, True)).float().cuda()
    Do = torch.from_numpy(_block_D(out, R, True)).float().cuda()
    with torch.no_grad():
        a = tp(x @ D1.T, yr)
        b = tp(x, y) @ Do.T
    err = ((a - b).abs().max() / b.abs().max()).item()
    assert err < 2e-5, err
    # weight gradient is additive over row blocks
    g = torch.randn(rows, tp.iro.dim, device="cuda")

    def gw(lo, hi):
        tp.weight.grad = None
        tp(x[lo:hi], y[lo:hi]).backward(g[lo:hi])
        return tp.weight.grad.double().clone()
    whole = gw(0, rows)
    parts = gw(0, 100_001) + gw(100_001, rows)
    err = ((whole - parts).abs().max() / whole.abs().max()).item()
    assert err < 2e-5, err


def test_gathered_segments_match_dense():
    """forward_cat([(x, dst), (x, src), (extra, None)]) == forward(cat(x[dst], x[src], extra)), values and gradients
    (the gathered gradients are scatter-added inside the kernel)."""
    in1, lmax, out = CASES["message2"]
    tp = _module(in1, lmax, out)
    torch.manual_seed(3)
    nn_, rows, dh = 3000, 40_001, 64
    x = torch.randn(nn_, dh, device="cuda", requires_grad=True)
    extra = torch.randn(rows, 2, device="cuda", requires_grad=True)
    dst = torch.sort(torch.randint(0, nn_, (rows,), device="cuda"))[0].int()
    src = torch.randint(0, nn_, (rows,), device="cuda").int()
    y = torch.randn(rows, 9, device="cuda")
    g = torch.randn(rows, tp.iro.dim, device="cuda")
    a = tp(torch.cat([x[dst.long()], x[src.long()], extra], 1), y)
    a.backward(g)
    want = (a.detach().clone(), x.grad.clone(), extra.grad.clone(), tp.weight.grad.clone())
    x.grad = extra.grad = tp.weight.grad = None
    b = tp.forward_cat([(x, dst, True), (x, src), (extra, None)], y)     # dst sorted: run-combined adds; src: 16-byte reds
    b.backward(g)
    assert torch.equal(b, want[0])
    for got, ref, what in ((x.grad, want[1], "x"), (extra.grad, want[2], "extra"), (tp.weight.grad, want[3], "weight")):
        err = ((got - ref).abs().max() / ref.abs().max()).item()
        assert err < 2e-5, (what, err)
