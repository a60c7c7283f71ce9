"""GPU parity of the fused l<=1 tensor-product kernels (through the C ABI) against
(1) the golden vectors produced by the unmodified reference, and
(2) the numpy oracle on seeded inputs at sizes spanning many tiles / CTAs,
including the fused options (gathered segments, gate, residual, sorted-segment sum).
Tolerance: 1e-5 relative to the largest reference magnitude, fp32 (north_star)."""
import os

import numpy as np
import pytest
import torch

from conftest import golden_l1tp_files, load_golden
from oracle import l1tp_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _close(a, ref, what, rtol=RTOL):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert a.shape == ref.shape, (what, a.shape, ref.shape)
    scale = max(np.abs(ref).max(), 1e-30)
    err = np.abs(a - ref).max() / scale
    assert err <= rtol, f"{what}: rel err {err:.3e} > {rtol}"


@pytest.mark.parametrize("path", golden_l1tp_files(), ids=lambda p: os.path.basename(p)[5:-4])
def test_module_matches_reference_golden(path):
    from se3gnn_b200.irreps import Irreps
    from models.segnn.l1_tensor_prod import L1TensorProduct
    rec = load_golden(path)
    meta = rec["meta"]
    tp = L1TensorProduct(Irreps(meta["in1"]), Irreps(meta["out"]), **meta["kwargs"])
    tp.load_state_dict({k[3:]: torch.from_numpy(rec[k]) for k in rec if k.startswith("sd_")})
    tp = tp.cuda()
    x = torch.from_numpy(rec["x"]).cuda().requires_grad_(True)
    y = torch.from_numpy(rec["y"]).cuda().requires_grad_(True)
    out = tp(x, y)
    assert out.is_contiguous() and out.dtype == torch.float32
    _close(out, rec["out_f64"], "out")
    out.backward(torch.from_numpy(rec["gout"]).cuda())
    _close(x.grad, rec["gx_f64"], "grad in1")
    _close(y.grad, rec["gy_f64"], "grad in2")
    for k, p in tp.named_parameters():
        _close(p.grad, rec[f"gw_{k}_f64"], f"grad {k}")


def _rand_weights(in1, out, seed):
    rng = np.random.default_rng(seed)
    i1, io = O.parse_irreps(in1), O.parse_irreps(out)
    w = {k: rng.uniform(-1, 1, s).astype(np.float32) for k, s in O.weight_shapes(i1, io).items()}
    a, _, _ = O.norm_factors(i1, io)
    nrm = O.norm_buffers(io, a)
    return w, nrm


def _to_lists(w, nrm, dev):
    ws, ns = [], []
    for s in ("l0e", "l0o", "l1e", "l1o"):
        ws.append(torch.from_numpy(w[f"weights_{s}"]).to(dev) if f"weights_{s}" in w else None)
        n = nrm[f"norm_{s}"]
        ns.append(torch.from_numpy(n.astype(np.float32)).to(dev) if n.size else None)
    return ws, ns


@pytest.mark.parametrize("in1,out,rows", [
    ("34x0e+10x1o+34x0e+10x1o+2x0e", "44x0e+10x1o", 20011),
    ("34x0e+10x1o", "44x0e+10x1o", 7000),
    ("4x0e+3x0o+2x1e+5x1o", "3x0e+2x0o+2x1e+3x1o", 3001),
    ("34x0e+10x1o", "1x1o", 1),
    ("96x0e+32x1o", "80x0e+24x1o", 513),
    # tile tails of the tcgen05 kernels (64-row tiles; 32-row tiles in the weight-gradient kernel)
    ("34x0e+10x1o+34x0e+10x1o+2x0e", "44x0e+10x1o", 65),
    ("34x0e+10x1o", "34x0e+10x1o", 63),
    ("34x0e+10x1o", "44x0e+10x1o", 1),
    ("2x1o+2x0e", "34x0e+10x1o", 130),
])
def test_plain_vs_oracle_large(in1, out, rows):
    from se3gnn_b200 import capi
    from se3gnn_b200.irreps import Irreps
    from se3gnn_b200.tp import TPConfig, get_plan, tp_layer
    dev = torch.device("cuda")
    w, nrm = _rand_weights(in1, out, 1)
    rng = np.random.default_rng(2)
    din, dout = O.irreps_dim(O.parse_irreps(in1)), O.irreps_dim(O.parse_irreps(out))
    x = rng.standard_normal((rows, din)).astype(np.float32)
    y = rng.standard_normal((rows, 4)).astype(np.float32)
    go = rng.standard_normal((rows, dout)).astype(np.float32)
    ws, ns = _to_lists(w, nrm, dev)
    for t in ws:
        if t is not None:
            t.requires_grad_(True)
    xt = torch.from_numpy(x).to(dev).requires_grad_(True)
    yt = torch.from_numpy(y).to(dev).requires_grad_(True)
    cfg = TPConfig(plan=get_plan(Irreps(in1), Irreps(out)), widths=(din,), need_gin2=True)
    tc0 = capi.tc_launch_count()
    o = tp_layer(cfg, rows, [xt], [None], yt, ws, ns)
    o.backward(torch.from_numpy(go).to(dev))
    if "0o" not in in1 and "1e" not in in1 and out != "1x1o" and "96x0e" not in in1 and din % 4 == 0:
        # a x0e + b x1o -> c x0e + d x1o: forward, weight-gradient and input-gradient all run on the tensor cores
        # (the weight-gradient kernel stages the cotangent rows with 16-byte cp.async: it declines dout % 4 != 0 here)
        # (the 8-column embedding input is declined by the tcgen05 weight-gradient kernel: generic fp32 kernel; the
        # first-generation tcgen05 kernel that used to take it was retired in round 2)
        want = (3 if dout % 4 == 0 else 2) - (1 if in1 == "2x1o+2x0e" else 0)
        assert capi.tc_launch_count() - tc0 == want, "tcgen05 kernels did not launch"
    w64 = {k: v.astype(np.float64) for k, v in w.items()}
    ref = O.forward(x.astype(np.float64), y.astype(np.float64), w64, nrm, in1, out)
    gx, gy, gw = O.backward(x.astype(np.float64), y.astype(np.float64), go.astype(np.float64), w64, nrm, in1, out)
    _close(o, ref, "out")
    _close(xt.grad, gx, "gx")
    _close(yt.grad, gy, "gy")
    for i, s in enumerate(("l0e", "l0o", "l1e", "l1o")):
        if ws[i] is not None:
            _close(ws[i].grad, gw[f"weights_{s}"], f"gw {s}", rtol=3e-5)  # long fp32 reductions over rows


def _gate_np(raw, ns, nv, cs, cg):
    sig = lambda z: 1 / (1 + np.exp(-z))
    s = raw[:, :ns]
    g = raw[:, ns:ns + nv]
    v = raw[:, ns + nv:].reshape(-1, nv, 3)
    return np.concatenate([cs * s * sig(s), (cg * sig(g)[:, :, None] * v).reshape(len(raw), -1)], axis=1)


def test_fused_message_layer_vs_oracle():
    """gather(x[dst]), gather(x[src]), extras -> TP -> gate  ;  TP -> gate -> segment-sum, with backward."""
    from se3gnn_b200 import capi
    from se3gnn_b200.gate import SIGMOID_CST, SILU_CST
    from se3gnn_b200.irreps import Irreps
    from se3gnn_b200.tp import TPConfig, get_plan, tp_layer
    dev = torch.device("cuda")
    rng = np.random.default_rng(5)
    N, E = 700, 9000
    hid = "34x0e+10x1o"
    in1 = f"{hid}+{hid}+2x0e"
    out = "44x0e+10x1o"
    dst = np.sort(rng.integers(0, N, E)).astype(np.int32)
    src = rng.integers(0, N, E).astype(np.int32)
    x = rng.standard_normal((N, 64)).astype(np.float32)
    ex = rng.standard_normal((E, 2)).astype(np.float32)
    y = rng.standard_normal((E, 4)).astype(np.float32)
    w1, n1 = _rand_weights(in1, out, 11)
    w2, n2 = _rand_weights(hid, out, 12)
    gagg = rng.standard_normal((N, 64)).astype(np.float32)

    # ---- oracle (fp64)
    f64 = lambda a: a.astype(np.float64)
    cat = np.concatenate([x[dst], x[src], ex], axis=1)
    w1d = {k: f64(v) for k, v in w1.items()}
    w2d = {k: f64(v) for k, v in w2.items()}
    raw1 = O.forward(f64(cat), f64(y), w1d, n1, in1, out)
    m1 = _gate_np(raw1, 34, 10, SILU_CST, SIGMOID_CST)
    raw2 = O.forward(m1, f64(y), w2d, n2, hid, out)
    m2 = _gate_np(raw2, 34, 10, SILU_CST, SIGMOID_CST)
    agg = np.zeros((N, 64))
    np.add.at(agg, dst, m2)

    # ---- CUDA
    xt = torch.from_numpy(x).to(dev).requires_grad_(True)
    ext = torch.from_numpy(ex).to(dev)
    yt = torch.from_numpy(y).to(dev)
    dt, st = torch.from_numpy(dst).to(dev), torch.from_numpy(src).to(dev)
    ws1, ns1 = _to_lists(w1, n1, dev)
    ws2, ns2 = _to_lists(w2, n2, dev)
    for t in ws1 + ws2:
        if t is not None:
            t.requires_grad_(True)
    cfg1 = TPConfig(plan=get_plan(Irreps(in1), Irreps(out)), widths=(64, 64, 2), epilogue=capi.EPI_GATE, gate_ns=34,
                    gate_cs=SILU_CST, gate_cg=SIGMOID_CST,
                    grad_modes=(capi.GRAD_SORTED, capi.GRAD_ATOMIC, capi.GRAD_NONE), share_grad={1: 0})
    cfg2 = TPConfig(plan=get_plan(Irreps(hid), Irreps(out)), widths=(64,), epilogue=capi.EPI_GATE, gate_ns=34,
                    gate_cs=SILU_CST, gate_cg=SIGMOID_CST, num_segments=N)
    tc0 = capi.tc_launch_count()
    m1t = tp_layer(cfg1, E, [xt, xt, ext], [dt, st, None], yt, ws1, ns1)
    assert capi.tc_launch_count() == tc0 + 1, "the SEGNN message TP must run on the tcgen05 path"
    _close(m1t, m1, "m1")
    aggt = tp_layer(cfg2, E, [m1t], [None], yt, ws2, ns2, seg_idx=dt)
    _close(aggt, agg, "agg")
    aggt.backward(torch.from_numpy(gagg).to(dev))

    # ---- oracle backward by finite composition: use torch fp64 autograd over the numpy forward restated in torch
    xd = torch.from_numpy(f64(x)).requires_grad_(True)
    W1 = {k: torch.from_numpy(v).requires_grad_(True) for k, v in w1d.items()}
    W2 = {k: torch.from_numpy(v).requires_grad_(True) for k, v in w2d.items()}

    def tp_torch(inp, yy, W, nrm, ir_in, ir_out):
        ci, co = O.species_columns(O.parse_irreps(ir_in)), O.species_columns(O.parse_irreps(ir_out))
        y0, y1 = yy[:, 0:1], yy[:, None, 1:4]
        s0e = inp[:, ci["0e"]]
        v1o = torch.stack([inp[:, ci["1o"] + c] for c in range(3)], -1)
        f0 = torch.cat([s0e * y0, O.C3 * (v1o * y1).sum(-1)], 1)
        f1 = torch.cat([O.C3 * s0e[:, :, None] * y1, O.C3 * v1o * y0[:, :, None]], 1)
        o0 = (f0 @ W["weights_l0e"]) * torch.from_numpy(nrm["norm_l0e"])
        o1 = torch.einsum("ekc,km->emc", f1, W["weights_l1o"]) * torch.from_numpy(nrm["norm_l1o"]).reshape(1, -1, 3)
        return torch.cat([o0, o1.reshape(len(inp), -1)], 1)  # out irreps are 0e block then 1o block

    def gate_t(raw):
        s, g, v = raw[:, :34], raw[:, 34:44], raw[:, 44:].reshape(-1, 10, 3)
        return torch.cat([SILU_CST * s * torch.sigmoid(s), (SIGMOID_CST * torch.sigmoid(g)[:, :, None] * v).reshape(len(raw), -1)], 1)

    yd = torch.from_numpy(f64(y))
    dl, sl = torch.from_numpy(dst).long(), torch.from_numpy(src).long()
    c = torch.cat([xd[dl], xd[sl], torch.from_numpy(f64(ex))], 1)
    m1r = gate_t(tp_torch(c, yd, W1, n1, in1, out))
    np.testing.assert_allclose(m1r.detach().numpy(), m1, rtol=1e-10, atol=1e-10)
    m2r = gate_t(tp_torch(m1r, yd, W2, n2, hid, out))
    aggr = torch.zeros(N, 64, dtype=torch.float64).index_add(0, dl, m2r)
    aggr.backward(torch.from_numpy(f64(gagg)))
    _close(xt.grad, xd.grad.numpy(), "grad x", rtol=3e-5)
    for i, s in enumerate(("l0e", "l0o", "l1e", "l1o")):
        if ws1[i] is not None:
            _close(ws1[i].grad, W1[f"weights_{s}"].grad.numpy(), f"gw1 {s}", rtol=3e-5)
            _close(ws2[i].grad, W2[f"weights_{s}"].grad.numpy(), f"gw2 {s}", rtol=3e-5)


def test_residual_epilogue_and_two_direct_segments():
    from se3gnn_b200.irreps import Irreps
    from se3gnn_b200.tp import TPConfig, get_plan, tp_layer
    dev = torch.device("cuda")
    rng = np.random.default_rng(9)
    N = 1234
    hid = "34x0e+10x1o"
    in1 = f"{hid}+{hid}"
    w, nrm = _rand_weights(in1, hid, 3)
    xa = rng.standard_normal((N, 64)).astype(np.float32)
    xb = rng.standard_normal((N, 64)).astype(np.float32)
    y = rng.standard_normal((N, 4)).astype(np.float32)
    go = rng.standard_normal((N, 64)).astype(np.float32)
    ws, ns = _to_lists(w, nrm, dev)
    a = torch.from_numpy(xa).to(dev).requires_grad_(True)
    b = torch.from_numpy(xb).to(dev).requires_grad_(True)
    cfg = TPConfig(plan=get_plan(Irreps(in1), Irreps(hid)), widths=(64, 64))
    o = tp_layer(cfg, N, [a, b], [None, None], torch.from_numpy(y).to(dev), ws, ns, resid=a)
    o.backward(torch.from_numpy(go).to(dev))
    f64 = lambda t: t.astype(np.float64)
    w64 = {k: f64(v) for k, v in w.items()}
    cat = np.concatenate([xa, xb], 1)
    ref = O.forward(f64(cat), f64(y), w64, nrm, in1, hid) + xa
    gx, _, _ = O.backward(f64(cat), f64(y), f64(go), w64, nrm, in1, hid)
    _close(o, ref, "out+resid")
    _close(a.grad, gx[:, :64] + go, "grad a (incl. residual)")
    _close(b.grad, gx[:, 64:], "grad b")


def test_zero_rows_and_errors():
    from se3gnn_b200.irreps import Irreps
    from models.segnn.l1_tensor_prod import L1TensorProduct
    tp = L1TensorProduct(Irreps("8x0e+4x1o")).cuda()
    o = tp(torch.zeros(0, 20, device="cuda"), torch.zeros(0, 4, device="cuda"))
    assert o.shape == (0, 20)
    with pytest.raises(RuntimeError):
        tp(torch.zeros(3, 20, device="cuda"), torch.zeros(2, 4, device="cuda"))


def test_equivariance():
    """Rotating inputs (scalars fixed, vectors and Y1 rotated) rotates the vector outputs."""
    from se3gnn_b200.irreps import Irreps
    from models.segnn.l1_tensor_prod import L1TensorProduct
    torch.manual_seed(0)
    tp = L1TensorProduct(Irreps("16x0e+8x1o")).cuda()
    x = torch.randn(300, 40, device="cuda")
    y = torch.randn(300, 4, device="cuda")
    q, _ = torch.linalg.qr(torch.randn(3, 3, dtype=torch.float64))
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    R = q.float().cuda()

    def rot(t, ns):
        s, v = t[:, :ns], t[:, ns:].reshape(len(t), -1, 3)
        return torch.cat([s, (v @ R.T).reshape(len(t), -1)], 1)

    o1 = rot(tp(x, y), 16)
    o2 = tp(rot(x, 16), rot(y, 1))
    assert (o1 - o2).abs().max() <= 2e-5 * o1.abs().max()
