#!/usr/bin/env python
"""One training step of a 1-layer SEGNN on a 100k-particle cloud (for ncu: few launches, realistic sizes)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scalable-e3-gnn_b200")):
    sys.path.insert(0, p)
import torch
from se3gnn_b200.pipeline import TrainStep, synthetic_cloud
from models.segnn.segnn import SEGNN

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 1
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = SEGNN(num_layers=layers).to(dev)
ts = TrainStep(model)
devt = [torch.from_numpy(x).to(dev) for x in synthetic_cloud(n, "plummer", seed=1)]
for _ in range(steps):
    loss = ts.step_device(*devt)
torch.cuda.synchronize()
print("loss", float(loss), "edges", ts.last_graph.e)
