#!/usr/bin/env python
"""ONE training step of the benchmarked workload (BASELINE configs[1]) between cudaProfilerStart / Stop, after two
untimed warm-up steps: the target of the ncu launch list committed under profiles/
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \\
        python tools/one_step.py [--particles 100000]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scalable-e3-gnn_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--particles", type=int, default=100_000)
ap.add_argument("--layers", type=int, default=4)
a = ap.parse_args()
from models.segnn.segnn import SEGNN  # noqa: E402
from se3gnn_b200 import capi  # noqa: E402
from se3gnn_b200.pipeline import TrainStep, synthetic_cloud  # noqa: E402

data = [torch.from_numpy(x).cuda() for x in synthetic_cloud(a.particles, "plummer", 1)]
torch.manual_seed(0)
ts = TrainStep(SEGNN(num_layers=a.layers).cuda())
for _ in range(2):
    ts.step_device(*data)
torch.cuda.synchronize()
n0 = capi.launch_count()
torch.cuda.profiler.start()
loss = ts.step_device(*data)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"one step: loss {float(loss):.6f}, {capi.launch_count() - n0} library launches, {ts.last_graph.e} edges")
