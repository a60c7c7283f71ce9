"""End-to-end parity of the hot path: GPU octree graph + 4-layer SEGNN (l_max=1) forward/backward
vs the CPU oracle (fp64 torch port of the reference TP in the public SEGNN layout) on the same
inputs and the same weights.  BASELINE config 0 size (1k particles)."""
import numpy as np
import pytest
import torch

from oracle import octree_oracle as T
from oracle.segnn_oracle import SEGNNOracle, graph_features

pytestmark = pytest.mark.gpu


def _relerr(a, ref):
    a = a.detach().cpu().double().numpy()
    ref = ref.detach().double().numpy()
    return np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-30)


@pytest.mark.parametrize("n,layers", [(1000, 4), (3000, 2)])
def test_segnn_forward_backward_vs_oracle(n, layers):
    from models.segnn.segnn import SEGNN
    from se3gnn_b200.octree import build_octree_graph
    rng = np.random.default_rng(0)
    pos = rng.standard_normal((n, 3)).astype(np.float32)
    vel = rng.standard_normal((n, 3)).astype(np.float32)
    mass = np.full(n, 1.0 / n, np.float32)
    target = rng.standard_normal((n, 3)).astype(np.float32)

    torch.manual_seed(0)
    model = SEGNN(num_layers=layers).cuda()
    g = build_octree_graph(torch.from_numpy(pos).cuda(), torch.from_numpy(vel).cuda(), torch.from_numpy(mass).cuda())
    out = model.forward_graph(g)
    loss = (out[:n] - torch.from_numpy(target).cuda()).square().mean()
    loss.backward()

    # oracle on the same graph (bit-exact by test_octree_gpu) and the GPU's own fp32 features
    ref_g = T.build_graph(pos)
    assert (ref_g["m"], len(ref_g["col"])) == (g.m, g.e)
    feats = graph_features(ref_g, pos, vel, mass)
    np.testing.assert_allclose(g.x_in.cpu().numpy(), feats["x_in"], rtol=1e-4, atol=1e-5)
    oracle = SEGNNOracle(num_layers=layers).double()
    oracle.load_state_dict({k: v.detach().cpu().double() for k, v in model.state_dict().items()})
    f64 = lambda t: t.detach().cpu().double()
    o_ref = oracle(f64(g.x_in), f64(g.node_attr), f64(g.edge_attr), f64(g.edge_extra), g.dst.cpu(), g.col.cpu())
    l_ref = (o_ref[:n] - torch.from_numpy(target).double()).square().mean()
    l_ref.backward()

    e_out = _relerr(out, o_ref)
    assert e_out <= 1e-5, f"node outputs rel err {e_out:.2e}"
    assert abs(loss.item() - l_ref.item()) <= 1e-5 * abs(l_ref.item())
    worst = 0.0
    pr = dict(oracle.named_parameters())
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        worst = max(worst, _relerr(p.grad, pr[k].grad))
    assert worst <= 5e-5, f"weight grads rel err {worst:.2e}"
    print(f"n={n} layers={layers}: out rel err {e_out:.2e}, worst grad rel err {worst:.2e}")
