"""The factorisation behind csrc/o3msg.cu (message product of the l <= 2 model by linearity: node tables, then per-edge
coupling) restated on the CPU (oracle/o3msg_oracle.py) equals the tensor product on the concatenated row
(oracle/lmax2_oracle.forward) to fp64 rounding, forward and every gradient."""
import pytest
import torch

from oracle import lmax2_oracle as O2
from oracle import o3msg_oracle as OM


@pytest.mark.parametrize("hidden,extras", [("23x0e+7x1o+4x2e", "2x0e"), ("5x0e+2x1o", "2x0e"), ("3x0e+2x1o+1x1e+2x2e", "1x0e+2x0e")])
def test_tables_equal_concatenated_product(hidden, extras):
    from oracle.l1tp_oracle import parse_irreps
    H, X = parse_irreps(hidden), parse_irreps(extras)
    gates = sum(m for m, l, _ in H if l > 0)
    out = [(sum(m for m, l, _ in H if l == 0) + gates, 0, 1)] + [(m, l, p) for m, l, p in H if l > 0]
    in2 = O2.sh_irreps(2)
    in1 = H + H + X
    g = torch.Generator().manual_seed(len(hidden))
    n, e = 13, 90
    dst = torch.randint(0, n, (e,), generator=g).sort().values
    src = torch.randint(0, n, (e,), generator=g)
    dh, dx = sum(m * (2 * l + 1) for m, l, _ in H), sum(m for m, _, _ in X)
    x = torch.randn(n, dh, generator=g, dtype=torch.float64, requires_grad=True)
    ex = torch.randn(e, dx, generator=g, dtype=torch.float64)
    y = torch.randn(e, 9, generator=g, dtype=torch.float64)
    ws = [torch.randn(s, generator=g, dtype=torch.float64, requires_grad=True) for s in O2.weight_shapes(in1, in2, out)]
    want = O2.forward(torch.cat([x[dst], x[src], ex], 1), y, ws, in1, in2, out)
    cot = torch.randn(want.shape, generator=g, dtype=torch.float64)
    gw = torch.autograd.grad((want * cot).sum(), [x] + ws)
    got = OM.message(x, dst, src, ex, y, ws, H, X, in2, out)
    gg = torch.autograd.grad((got * cot).sum(), [x] + ws)
    assert (got - want).abs().max() < 1e-12 * want.abs().max()
    for a, b in zip(gg, gw):
        assert (a - b).abs().max() <= 1e-12 * max(1.0, b.abs().max().item())
