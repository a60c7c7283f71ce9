"""SEGNN l_max = 2 (BASELINE configs[2] as a parity case): the CUDA model (`models/segnn/segnn_l2.py`, every tensor
product on the l <= 2 kernels, graph + SH(2) attributes from the GPU builder) against the fp64 CPU specification
`oracle/segnn_l2_oracle.py` on the same graph and weights: node outputs within 1e-5 of the largest reference magnitude
(fp32, north_star), weight gradients within 5e-4 (sums over every edge and layer in fp32, with atomics in the torch
aggregation: run-to-run noise of the same size, cf. tools/check_dd.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cloud(n, seed):
    rng = np.random.default_rng(seed)
    u = rng.random(n)
    r = np.minimum(1.0 / np.sqrt(u ** (-2.0 / 3.0) - 1.0), 10.0)
    d = rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return (r[:, None] * d).astype(np.float32), rng.standard_normal((n, 3)).astype(np.float32)


@pytest.mark.parametrize("n,hidden,layers", [(300, "23x0e+7x1o+4x2e", 2), (700, "8x0e+3x1o+2x2e", 4)])
def test_model_matches_oracle(n, hidden, layers):
    from models.segnn.segnn_l2 import SEGNNL2
    from oracle.segnn_l2_oracle import SEGNNL2Oracle
    from se3gnn_b200 import capi
    from se3gnn_b200.octree import build_octree_graph, sh2_attributes
    pos, vel = _cloud(n, n)
    g = build_octree_graph(torch.from_numpy(pos).cuda(), torch.from_numpy(vel).cuda(), leaf_size=16)
    ea, na = sh2_attributes(g)
    torch.manual_seed(0)
    ref = SEGNNL2Oracle(hidden, layers)
    model = SEGNNL2(hidden, layers).cuda()
    model.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    n0 = capi.launch_count()
    out = model(g.x_in, na, ea, g.edge_extra, g.dst, g.col)
    cot = torch.randn(out.shape, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    (out * cot.float().cuda()).sum().backward()
    torch.cuda.synchronize()
    ntp = 2 + 4 * layers + 1
    assert capi.launch_count() - n0 >= 3 * ntp - 1        # every tensor product ran on this library's kernels
    want = ref(g.x_in.cpu().double(), na.cpu().double(), ea.cpu().double(), g.edge_extra.cpu().double(), g.dst.cpu(), g.col.cpu())
    (want * cot).sum().backward()
    err = (out.detach().cpu().double() - want.detach()).abs().max() / want.detach().abs().max()
    assert err < 1e-5, f"node outputs: rel err {err:.3e}"
    for (k, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
        e = (p.grad.cpu().double() - q.grad).abs().max() / max(q.grad.abs().max().item(), 1e-30)
        assert e < 5e-4, f"grad {k}: rel err {e:.3e}"
