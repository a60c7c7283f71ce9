"""Node-table product of the l_max = 2 message layer (64 x 1 -> 446) forward/backward on 1.1M rows: launch list probe."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scalable-e3-gnn_b200"))
import torch
import __graft_entry__ as ge
ge.build()
from se3gnn_b200.irreps import Irreps
from se3gnn_b200.o3tp import O3TensorProduct
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_122_695
tp = O3TensorProduct(Irreps("23x0e+7x1o+4x2e"), Irreps("45x0e+52x1o+49x2e"), Irreps("1x0e")).cuda()
print("plan: split", tp._plan.split_backward, "tc", tp._plan.tc_weight_grad, "tiles", tp._plan.tile_fwd, tp._plan.tile_bwd)
x = torch.randn(rows, 64, device="cuda", requires_grad=True)
y = torch.ones(rows, 1, device="cuda")
g = torch.randn(rows, 446, device="cuda")
for _ in range(3):
    out = tp(x, y)
    out.backward(g)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record(); out = tp(x, y); e[1].record(); out.backward(g); e[2].record()
torch.cuda.synchronize()
print("fwd ms", e[0].elapsed_time(e[1]), "bwd ms", e[1].elapsed_time(e[2]))
