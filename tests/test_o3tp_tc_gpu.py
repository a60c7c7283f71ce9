"""The tensor-core weight gradient of the l <= 2 tensor product (`o3tp_tc_gw_kernel`, csrc/o3tp_tc_gw.cu) and the
linear-map input gradient of scalar-second-input products (csrc/o3tp_lin.cu) against the fp64 specification
oracle/lmax2_oracle.py: the shapes of the l_max = 2 model (message 2, update 2, node tables), row counts around the
32-row tile (the remainder goes through the SIMT kernel), and a misaligned second input (the product must take the
SIMT path and give the same numbers).  1e-5 of the largest reference magnitude, fp32 (3xTF32 on the tensor cores)."""
import numpy as np
import pytest
import torch

from oracle import lmax2_oracle as O2

pytestmark = pytest.mark.gpu

H = [(23, 0, 1), (7, 1, -1), (4, 2, 1)]
SHAPES = {
    "message2": (H, O2.sh_irreps(2), [(34, 0, 1), (7, 1, -1), (4, 2, 1)]),
    "update2": (H, O2.sh_irreps(2), H),
    "node_tables": (H, [(1, 0, 1)], [(45, 0, 1), (52, 1, -1), (49, 2, 1)]),
    "readout": (H, O2.sh_irreps(2), [(1, 1, -1)]),
}


def _spec(ir):
    return "+".join(f"{m}x{l}{'e' if p == 1 else 'o'}" for m, l, p in ir)


def _run(name, rows, y_requires_grad=False, misalign=False):
    from se3gnn_b200 import capi
    from se3gnn_b200.irreps import Irreps
    from se3gnn_b200.o3tp import O3TensorProduct
    in1, in2, out = SHAPES[name]
    torch.manual_seed(rows)
    tp = O3TensorProduct(Irreps(_spec(in1)), Irreps(_spec(out)), Irreps(_spec(in2))).cuda()
    rng = np.random.default_rng(rows + len(name))
    x = rng.standard_normal((rows, tp.in1_dim))
    y = rng.standard_normal((rows, tp.in2_dim))
    g = rng.standard_normal((rows, tp.iro.dim))
    ws, o = [], 0
    w = tp.weight.detach().cpu().double()
    for shp in O2.weight_shapes(in1, in2, out):
        ws.append(w[o:o + shp[0] * shp[1]].reshape(shp).clone().requires_grad_())
        o += shp[0] * shp[1]
    xt = torch.from_numpy(x).requires_grad_()
    yt = torch.from_numpy(y).requires_grad_()
    ref = O2.forward(xt, yt, ws, in1, in2, out)
    ref.backward(torch.from_numpy(g))
    xg = torch.from_numpy(x).float().cuda().requires_grad_()
    if misalign:   # a view whose data pointer is 4 bytes past a 16-byte boundary
        buf = torch.zeros(rows * tp.in2_dim + 1, device="cuda")
        yg = buf[1:].view(rows, tp.in2_dim)
        yg.copy_(torch.from_numpy(y).float())
        assert yg.data_ptr() % 16 != 0
    else:
        yg = torch.from_numpy(y).float().cuda()
    yg.requires_grad_(y_requires_grad)
    t0 = capi.tc_launch_count()
    res = tp(xg, yg)
    res.backward(torch.from_numpy(g).float().cuda())
    torch.cuda.synchronize()
    used_tc = capi.tc_launch_count() - t0

    def rel(a, b):
        return float((a.detach().cpu().double() - b).abs().max() / b.abs().max().clamp_min(1e-30))
    assert rel(res, ref.detach()) < 1e-5
    assert rel(xg.grad, xt.grad) < 1e-5
    assert rel(tp.weight.grad, torch.cat([v.grad.reshape(-1) for v in ws])) < 1e-5
    if y_requires_grad:
        assert rel(yg.grad, yt.grad) < 1e-5
    return tp, used_tc


@pytest.mark.parametrize("name", list(SHAPES))
@pytest.mark.parametrize("rows", [31, 32, 33, 2049, 20000])
def test_model_shapes_against_oracle(name, rows):
    tp, used_tc = _run(name, rows)
    assert tp._plan.tc_weight_grad, "these shapes have a dense in1 of 64 columns: the tcgen05 kernel must cover them"
    assert used_tc == (1 if rows >= 32 else 0)
    assert tp._plan.linear_maps == (name == "node_tables")


def test_second_input_gradient_and_misaligned_rows():
    _, used = _run("message2", 4099, y_requires_grad=True)
    assert used == 1
    _, used = _run("message2", 4099, misalign=True)     # cp.async.bulk needs 16-byte aligned rows: SIMT path, same numbers
    assert used == 0
    _, used = _run("node_tables", 3000, y_requires_grad=True)   # gradient of the scalar input: general kernels
    assert used in (0, 1)
