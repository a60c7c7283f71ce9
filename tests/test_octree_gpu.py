"""GPU octree/graph builder vs the CPU specification: cell assignment and the canonically
sorted edge list must be bit-exact (integer work); float node/edge data within 1e-5."""
import numpy as np
import pytest
import torch

from oracle import octree_oracle as T

pytestmark = pytest.mark.gpu


def _plummer(n, seed):
    rng = np.random.default_rng(seed)
    u = rng.random(n)
    r = 1.0 / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
    r = np.minimum(r, 10.0)
    d = rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return (r[:, None] * d).astype(np.float32)


def _cases():
    rng = np.random.default_rng(0)
    yield "uniform_1k", rng.random((1000, 3)).astype(np.float32), 32
    yield "uniform_100k", rng.random((100000, 3)).astype(np.float32), 32
    yield "plummer_50k", _plummer(50000, 1), 32
    yield "plummer_leaf8", _plummer(20000, 2), 8
    p = rng.random((3000, 3)).astype(np.float32)
    p[:200] = p[0]
    yield "duplicates", p, 32
    yield "tiny", rng.random((5, 3)).astype(np.float32), 32
    yield "single", np.zeros((1, 3), np.float32), 32
    yield "ragged_2049", rng.random((2049, 3)).astype(np.float32), 32


@pytest.mark.parametrize("name,pos,leaf", list(_cases()), ids=[c[0] for c in _cases()])
def test_graph_bit_exact(name, pos, leaf):
    from se3gnn_b200.octree import build_octree_graph
    ref = T.build_graph(pos, leaf_size=leaf)
    g = build_octree_graph(torch.from_numpy(pos).cuda(), leaf_size=leaf, features=False)
    assert (g.n, g.m, g.e) == (ref["n"], ref["m"], len(ref["col"]))
    eq = lambda a, b: np.testing.assert_array_equal(a.cpu().numpy(), b)
    eq(g.keys, ref["keys"].view(np.int64))
    eq(g.order, ref["order"])
    for k in ("cell_start", "cell_count", "cell_level", "cell_parent", "cell_first_child", "cell_nchild"):
        eq(getattr(g, k), ref[k])
    eq(g.cell_key, ref["cell_key"].view(np.int64))
    eq(g.leaf_of_rank, ref["leaf_of_rank"])
    eq(g.cell_of_particle, ref["cell_of_particle"])
    eq(g.rowptr, ref["rowptr"])
    eq(g.col, ref["col"])
    eq(g.dst, ref["dst"])
    nl = len(ref["level_ptr"]) - 1
    assert g.nlevels == nl
    eq(g.level_ptr[:nl + 1], ref["level_ptr"])


def test_node_and_edge_features():
    from se3gnn_b200.octree import build_octree_graph
    rng = np.random.default_rng(4)
    n = 5000
    pos = _plummer(n, 7)
    vel = rng.standard_normal((n, 3)).astype(np.float32)
    mass = (rng.random(n).astype(np.float32) + 0.5) / n
    ref = T.build_graph(pos)
    g = build_octree_graph(torch.from_numpy(pos).cuda(), torch.from_numpy(vel).cuda(), torch.from_numpy(mass).cuda())
    mm, com, cv = T.cell_moments(ref, pos, vel, mass)
    o = ref["order"]
    npos = np.concatenate([pos[o].astype(np.float64), com])
    nvel = np.concatenate([vel[o].astype(np.float64), cv])
    nmass = np.concatenate([mass[o].astype(np.float64), mm])
    np.testing.assert_allclose(g.node_pos.cpu().numpy(), npos, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(g.node_vel.cpu().numpy(), nvel, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(g.node_mass.cpu().numpy(), nmass, rtol=1e-5)
    # edge geometry from the GPU's own fp32 node positions (so the comparison isolates the kernel)
    P = g.node_pos.cpu().numpy().astype(np.float64)
    M = g.node_mass.cpu().numpy().astype(np.float64)
    V = g.node_vel.cpu().numpy().astype(np.float64)
    d, s = ref["dst"], ref["col"]
    rel = P[s] - P[d]
    r = np.linalg.norm(rel, axis=1)
    sh0, sh1 = 0.28209479177387814, 0.4886025119029199
    unit = np.where(r[:, None] > 0, rel / np.maximum(r, 1e-300)[:, None], 0.0)
    ea = np.concatenate([np.full((len(d), 1), sh0), sh1 * unit], 1)
    got = g.edge_attr.cpu().numpy()
    ok = r > 1e-4  # direction of nearly coincident points is ill-conditioned in fp32
    np.testing.assert_allclose(got[ok], ea[ok], rtol=1e-4, atol=2e-4)
    ex = np.stack([r, (n * M[d]) * (n * M[s])], 1)
    np.testing.assert_allclose(g.edge_extra.cpu().numpy(), ex, rtol=1e-4, atol=1e-6)
    # node attr = mean incoming edge_attr (GPU's) + SH(vel)
    ea_gpu = got.astype(np.float64)
    na = np.zeros((len(P), 4))
    np.add.at(na, d, ea_gpu)
    deg = np.diff(ref["rowptr"])
    na /= np.maximum(deg, 1)[:, None]
    vn = np.linalg.norm(V, axis=1)
    na += np.concatenate([np.full((len(P), 1), sh0), sh1 * V / np.maximum(vn, 1e-300)[:, None]], 1)
    np.testing.assert_allclose(g.node_attr.cpu().numpy(), na, rtol=1e-4, atol=1e-5)
    xin = np.concatenate([P - P[n], V, vn[:, None], (M * n)[:, None]], 1)
    np.testing.assert_allclose(g.x_in.cpu().numpy(), xin, rtol=1e-5, atol=1e-6)
    assert g.edge_index.shape == (2, g.e)


def test_sh2_attributes():
    """SH(2) edge / node attributes for the l <= 2 tensor product against `oracle/lmax2_oracle.spherical_harmonics`
    (fp64) evaluated on the GPU's own fp32 node positions; the l <= 1 columns must be the l_max = 1 builder's."""
    from oracle import lmax2_oracle as O2
    from se3gnn_b200.octree import build_octree_graph, sh2_attributes
    rng = np.random.default_rng(5)
    n = 4000
    pos = _plummer(n, 9)
    pos[:5] = pos[0]                                          # coincident particles: zero-length edges, Y0 only
    vel = rng.standard_normal((n, 3)).astype(np.float32)
    vel[:7] = 0.0                                             # zero velocity: Y0 only
    g = build_octree_graph(torch.from_numpy(pos).cuda(), torch.from_numpy(vel).cuda())
    ea, na = sh2_attributes(g)
    assert ea.shape == (g.e, 9) and na.shape == (g.n + g.m, 9)
    assert torch.equal(ea[:, :4], g.edge_attr)
    P = g.node_pos.cpu().numpy().astype(np.float64)
    V = g.node_vel.cpu().numpy().astype(np.float64)
    d, s = g.dst.cpu().numpy(), g.col.cpu().numpy()
    rel = P[s] - P[d]
    want = O2.spherical_harmonics(rel, 2)
    ok = np.linalg.norm(rel, axis=1) > 1e-4
    got = ea.cpu().numpy()
    np.testing.assert_allclose(got[ok], want[ok], rtol=1e-4, atol=5e-4)
    zero_len = np.linalg.norm(rel, axis=1) == 0
    assert zero_len.any() and not got[zero_len][:, 1:].any()
    wn = np.zeros((len(P), 9))
    np.add.at(wn, d, got.astype(np.float64))
    wn /= np.maximum(np.diff(g.rowptr.cpu().numpy()), 1)[:, None]
    wn += O2.spherical_harmonics(V, 2)
    np.testing.assert_allclose(na.cpu().numpy(), wn, rtol=1e-4, atol=1e-5)
    # the attributes feed the l <= 2 tensor product
    from se3gnn_b200.irreps import Irreps
    from se3gnn_b200.o3tp import O3TensorProduct
    tp = O3TensorProduct(Irreps("8x0e+4x1o+2x2e"), Irreps("8x0e+4x1o+2x2e")).cuda()
    out = tp(torch.randn(g.e, tp.in1_dim, device="cuda"), ea)
    assert out.shape == (g.e, tp.iro.dim) and torch.isfinite(out).all()
