"""Cost of the decomposition bookkeeping on one GPU: global build at world x 100k particles + local_graph of rank 0."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scalable-e3-gnn_b200")):
    sys.path.insert(0, p)
import torch
from se3gnn_b200 import domain
from se3gnn_b200.octree import build_octree_graph
from se3gnn_b200.pipeline import synthetic_cloud
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
pos, vel, mass, tgt = [torch.from_numpy(x).to(dev) for x in synthetic_cloud(100_000 * world, "plummer", seed=1)]
def tm(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3, r
t_b, g = tm(lambda: build_octree_graph(pos, vel, mass))
t_l, lg = tm(lambda: domain.local_graph(0, world, g.n, g.cell_start, g.leaf_of_rank, g.dst, g.col, rowptr=g.rowptr))
t_s, _ = tm(lambda: (g.x_in.index_select(0, lg.own_ids), g.node_attr.index_select(0, lg.own_ids),
                     domain.take_edges(lg, g.edge_attr), domain.take_edges(lg, g.edge_extra)))
pos1 = pos[:100_000].contiguous(); vel1 = vel[:100_000].contiguous(); m1 = mass[:100_000].contiguous()
t_b1, _ = tm(lambda: build_octree_graph(pos1, vel1, m1))
print(f"world {world}: global build {t_b:.2f} ms (single-GPU-size build {t_b1:.2f} ms), local_graph {t_l:.2f} ms, slicing {t_s:.2f} ms; "
      f"edges global {g.e} local {lg.e} halo {lg.n_halo}")
