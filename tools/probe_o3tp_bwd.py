"""One forward/backward of the l <= 2 message-2 product (64 -> 75, SH(2)) on 2M rows: for the ncu launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scalable-e3-gnn_b200"))
import torch
import __graft_entry__ as ge
ge.build()
from se3gnn_b200.irreps import Irreps
from se3gnn_b200.o3tp import O3TensorProduct
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
tp = O3TensorProduct(Irreps("23x0e+7x1o+4x2e"), Irreps("34x0e+7x1o+4x2e"), Irreps.spherical_harmonics(2)).cuda()
x = torch.randn(rows, 64, device="cuda", requires_grad=True)
y = torch.randn(rows, 9, device="cuda")
g = torch.randn(rows, 75, device="cuda")
for _ in range(3):
    out = tp(x, y)
    out.backward(g)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record(); out = tp(x, y); e[1].record(); out.backward(g); e[2].record()
torch.cuda.synchronize()
print("fwd ms", e[0].elapsed_time(e[1]), "bwd ms", e[1].elapsed_time(e[2]))
