#!/usr/bin/env python
"""Host-side enqueue time of one training step vs its GPU time (is the step launch-bound?)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scalable-e3-gnn_b200")):
    sys.path.insert(0, p)
import torch
from se3gnn_b200.pipeline import TrainStep, synthetic_cloud
from se3gnn_b200.octree import build_octree_graph
from models.segnn.segnn import SEGNN
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = SEGNN(num_layers=4).to(dev)
ts = TrainStep(model)
devt = [torch.from_numpy(x).to(dev) for x in synthetic_cloud(100_000, "plummer", seed=1)]
for _ in range(3):
    ts.step_device(*devt)
torch.cuda.synchronize()
K = 10
t0 = time.perf_counter()
for _ in range(K):
    ts.step_device(*devt)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"enqueue {1e3*(t1-t0)/K:.2f} ms/step, total {1e3*(t2-t0)/K:.2f} ms/step")
# pieces
def tm(f, n=10):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): r = f()
    e = time.perf_counter(); torch.cuda.synchronize(); s = time.perf_counter()
    return 1e3*(e-t)/n, 1e3*(s-t)/n
print("build graph (enqueue, total)", tm(lambda: build_octree_graph(*devt[:3])))
g = build_octree_graph(*devt[:3])
print("forward", tm(lambda: model.forward_graph(g)))
def fb():
    out = model.forward_graph(g); out.square().mean().backward()
print("fwd+bwd", tm(fb))
