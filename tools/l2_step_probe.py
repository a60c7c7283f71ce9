"""One l_max = 2 training step at a given size (for ncu) + degree statistics of the graph."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scalable-e3-gnn_b200"))
import numpy as np, torch
import __graft_entry__ as ge
ge.build()
from models.segnn.segnn_l2 import SEGNNL2
from se3gnn_b200.octree import build_octree_graph, sh2_attributes
from se3gnn_b200.pipeline import synthetic_cloud
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
pos, vel, mass, target = (torch.from_numpy(a).cuda() for a in synthetic_cloud(n, "plummer", 1))
torch.manual_seed(0)
model = SEGNNL2("23x0e+7x1o+4x2e", 4).cuda()
for it in range(2):
    g = build_octree_graph(pos, vel, mass, leaf_size=32)
    out = model.forward_graph(g, sh2_attributes(g))
    loss = (out[:n] - target).square().mean()
    loss.backward()
torch.cuda.synchronize()
deg_in = torch.bincount(g.dst.long(), minlength=g.n + g.m)
deg_out = torch.bincount(g.col.long(), minlength=g.n + g.m)
for name, d in (("in", deg_in), ("out", deg_out)):
    q = torch.quantile(d.float(), torch.tensor([0.5, 0.9, 0.99, 0.999], device=d.device)).tolist()
    print(f"degree {name}: mean {d.float().mean():.1f} median/p90/p99/p99.9 {q} max {int(d.max())} ; cells only max {int(d[g.n:].max())} mean {d[g.n:].float().mean():.1f}")
print("loss", float(loss))
