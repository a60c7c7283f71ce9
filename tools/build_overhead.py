import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scalable-e3-gnn_b200")):
    sys.path.insert(0, p)
import torch, ctypes as C
from se3gnn_b200 import capi
from se3gnn_b200.pipeline import synthetic_cloud
from se3gnn_b200 import octree as O
dev = torch.device("cuda", 0)
devt = [torch.from_numpy(x).to(dev) for x in synthetic_cloud(100_000, "plummer", seed=1)]
lib = capi.lib()
acc = {}
class W:
    def __init__(s, lib): s.__dict__["_l"] = lib
    def __getattr__(s, name):
        f = getattr(s._l, name)
        def g(*a):
            t = time.perf_counter(); r = f(*a); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t; return r
        return g
capi_lib = capi.lib
capi.lib = lambda: W(lib)
for _ in range(3): O.build_octree_graph(*devt[:3])
torch.cuda.synchronize(); acc.clear()
K = 20
t0 = time.perf_counter()
for _ in range(K): g = O.build_octree_graph(*devt[:3])
torch.cuda.synchronize()
tot = (time.perf_counter() - t0) / K * 1e3
print(f"build_octree_graph: {tot:.3f} ms/call")
for k, v in acc.items(): print(f"   {k:24s} {v/K*1e3:.3f} ms")
capi.lib = capi_lib
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(K): g = O.build_octree_graph(*devt[:3])
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
