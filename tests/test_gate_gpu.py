"""The stand-alone gate kernels (`se3_gate_forward/backward`, csrc/gate.cu) against the defining formula in fp64
(public SEGNN Gate: normalize2mom silu on the scalars, normalize2mom sigmoid gates on every l > 0 channel)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(raw, ns, blocks):
    from se3gnn_b200.gate import SIGMOID_CST, SILU_CST
    ng = sum(c for c, _ in blocks)
    g = SIGMOID_CST * torch.sigmoid(raw[:, ns:ns + ng])
    out = [SILU_CST * torch.nn.functional.silu(raw[:, :ns])]
    o, k = ns + ng, 0
    for c, d in blocks:
        out.append((raw[:, o:o + c * d].reshape(-1, c, d) * g[:, k:k + c, None]).reshape(len(raw), -1))
        o += c * d
        k += c
    return torch.cat(out, 1)


@pytest.mark.parametrize("rows,ns,blocks", [(1000, 23, [(7, 3), (4, 5)]), (37, 5, [(2, 3)]), (513, 0, [(3, 5), (1, 3), (2, 1)]),
                                            (64, 9, []), (1, 1, [(1, 3), (1, 5), (1, 3), (1, 5)])])
def test_gate_matches_formula(rows, ns, blocks):
    from se3gnn_b200 import capi
    from se3gnn_b200.gate import irreps_gate
    d_raw = ns + sum(c for c, _ in blocks) + sum(c * d for c, d in blocks)
    torch.manual_seed(rows)
    raw = (2.0 * torch.randn(rows, d_raw, dtype=torch.float64)).requires_grad_()
    want = _ref(raw, ns, blocks)
    cot = torch.randn(want.shape, dtype=torch.float64)
    (want * cot).sum().backward()
    x = raw.detach().float().cuda().requires_grad_()
    n0 = capi.launch_count()
    got = irreps_gate(x, ns, blocks)
    (got * cot.float().cuda()).sum().backward()
    torch.cuda.synchronize()
    assert capi.launch_count() - n0 == 2
    assert got.shape == want.shape
    assert (got.detach().cpu().double() - want.detach()).abs().max() <= 1e-5 * max(1.0, want.detach().abs().max().item())
    assert (x.grad.cpu().double() - raw.grad).abs().max() <= 1e-5 * max(1.0, raw.grad.abs().max().item())
