"""Message product of the l <= 2 model by linearity (se3gnn_b200/o3msg.py, csrc/o3msg.cu: node tables + per-edge coupling)
against the tensor product on the concatenated row it replaces (`O3TensorProduct.forward_cat`, itself checked against
oracle/lmax2_oracle.py in test_o3tp_gpu.py), on random multigraphs with halo rows, isolated nodes and duplicate edges:
outputs within 1e-5, input gradients within 1e-5, weight gradients within 1e-4 of the largest reference magnitude (fp32
sums over all edges in a different order)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("hidden,n_dst,n_all,e", [("23x0e+7x1o+4x2e", 50, 60, 900), ("8x0e+3x1o+2x2e", 300, 300, 5000),
                                                  ("5x0e+2x1o", 7, 9, 40), ("6x0e+3x1o+1x1e", 40, 44, 600)])
def test_tables_equal_tensor_product(hidden, n_dst, n_all, e):
    from se3gnn_b200 import capi, o3msg
    from se3gnn_b200.irreps import Irreps
    from se3gnn_b200.msg import build_edge_index
    from se3gnn_b200.o3tp import O3TensorProduct
    g = torch.Generator().manual_seed(n_dst * 7 + e)
    h = Irreps(hidden)
    gates = sum(m.mul for m in h if m.ir.l > 0)
    out = Irreps(f"{h.count('0e') + gates}x0e+" + "+".join(f"{m.mul}x{m.ir}" for m in h if m.ir.l > 0))
    tp = O3TensorProduct(Irreps(f"{hidden}+{hidden}+2x0e"), out, Irreps.spherical_harmonics(2)).cuda()
    assert o3msg.supported(tp, hidden, "2x0e")
    dst = torch.randint(0, n_dst, (e,), generator=g).sort().values.int().cuda()
    dst[dst == 3] = 4                                   # node 3 has no incoming edge
    src = torch.randint(0, n_all, (e,), generator=g).int().cuda()
    xe = torch.randn(n_all, h.dim, generator=g).cuda().requires_grad_(True)
    y = torch.randn(e, 9, generator=g).cuda()
    ex = torch.randn(e, 2, generator=g).cuda()
    cot = torch.randn(e, out.dim, generator=g).cuda()
    want = tp.forward_cat([(xe, dst, True), (xe, src), (ex, None)], y)
    (want * cot).sum().backward()
    gx_w, gw_w = xe.grad.clone(), tp.weight.grad.clone()
    xe.grad = None
    tp.weight.grad = None
    ei = build_edge_index(dst, src, n_dst, n_all)
    n0 = capi.launch_count()
    got = o3msg.tables_for(tp, hidden, "2x0e")(tp.weight, xe, y, ex, ei)
    (got * cot).sum().backward()
    torch.cuda.synchronize()
    assert capi.launch_count() - n0 >= 7   # two node tables, edge forward, two transposed passes, two node backwards
    assert _rel(got.detach(), want.detach()) < 1e-5
    assert _rel(xe.grad, gx_w) < 1e-5
    assert _rel(tp.weight.grad, gw_w) < 1e-4


def test_unsupported_layouts_are_declined():
    from se3gnn_b200 import o3msg
    from se3gnn_b200.irreps import Irreps
    from se3gnn_b200.o3tp import O3TensorProduct
    sh = Irreps.spherical_harmonics(2)
    tp = O3TensorProduct(Irreps("4x0e+2x1o+4x0e+2x1o+1x1o"), Irreps("6x0e+2x1o"), sh).cuda()
    assert not o3msg.supported(tp, "4x0e+2x1o", "1x1o")          # vector extras
    tp = O3TensorProduct(Irreps("4x0e+2x1o+3x0e+2x1o+2x0e"), Irreps("6x0e+2x1o"), sh).cuda()
    assert not o3msg.supported(tp, "4x0e+2x1o", "2x0e")          # in1 is not hidden + hidden + extras


@pytest.mark.parametrize("mode", ["table", "tp"])
def test_model_modes_agree_with_oracle(mode, monkeypatch):
    """both message-1 paths of the l_max = 2 model against the fp64 oracle model (test_segnn_l2_gpu.py runs the default)"""
    import numpy as np
    from models.segnn.segnn_l2 import SEGNNL2
    from oracle.segnn_l2_oracle import SEGNNL2Oracle
    from se3gnn_b200.octree import build_octree_graph, sh2_attributes
    monkeypatch.setenv("SE3_L2_MSG", mode)
    rng = np.random.default_rng(5)
    pos = rng.standard_normal((400, 3)).astype(np.float32)
    vel = rng.standard_normal((400, 3)).astype(np.float32)
    g = build_octree_graph(torch.from_numpy(pos).cuda(), torch.from_numpy(vel).cuda(), leaf_size=16)
    ea, na = sh2_attributes(g)
    torch.manual_seed(0)
    ref = SEGNNL2Oracle("8x0e+3x1o+2x2e", 2)
    model = SEGNNL2("8x0e+3x1o+2x2e", 2).cuda()
    model.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    out = model(g.x_in, na, ea, g.edge_extra, g.dst, g.col)
    cot = torch.randn(out.shape, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    (out * cot.float().cuda()).sum().backward()
    want = ref(g.x_in.cpu().double(), na.cpu().double(), ea.cpu().double(), g.edge_extra.cpu().double(), g.dst.cpu(), g.col.cpu())
    (want * cot).sum().backward()
    assert _rel(out.detach().cpu().double(), want.detach()) < 1e-5
    for (k, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        assert _rel(p.grad.cpu().double(), q.grad) < 5e-4, k
