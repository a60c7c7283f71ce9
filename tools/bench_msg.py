#!/usr/bin/env python
"""Message layer alone (graph of an n-particle cloud, one layer, forward + backward), CUDA-event timed per library call.
    python tools/bench_msg.py [--particles 100000] [--reps 5] [--mode fused|table|tp]
Used under ncu for the per-kernel captures in profiles/ (the bench line itself comes from bench.py)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scalable-e3-gnn_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--particles", type=int, default=100_000)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--kind", default="plummer")
ap.add_argument("--mode", default="fused")
ap.add_argument("--timing", action="store_true", help="print the phase cycle counters of the fused forward kernel")
a = ap.parse_args()
os.environ["SE3_MSG"] = a.mode

from se3gnn_b200 import capi  # noqa: E402
from se3gnn_b200.octree import build_octree_graph  # noqa: E402
from se3gnn_b200.pipeline import synthetic_cloud  # noqa: E402
from models.segnn.segnn import SEGNN  # noqa: E402

pos, vel, mass, target = (torch.from_numpy(x).cuda() for x in synthetic_cloud(a.particles, a.kind, 1))
torch.manual_seed(0)
model = SEGNN(num_layers=1).cuda()
g = build_octree_graph(pos, vel, mass)


def step():
    out = model.forward_graph(g)
    out.square().mean().backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
capi.profile_begin()
for _ in range(a.reps):
    step()
prof = capi.profile_end()
agg = {}
for tag, ms, nb, fl in prof:
    r = agg.setdefault(tag, [0.0, 0.0, 0])
    r[0] += ms
    r[1] += nb
    r[2] += 1
if a.timing:
    from se3gnn_b200 import msg as _msg
    _msg.DBG_TIMING = []
    step()
    torch.cuda.synchronize()
    t = _msg.DBG_TIMING[0].double()            # [148][2][8]
    names = ["build", "bar2 wait", "drain(+issue)", "bar1 wait", "finish+prefetch", "tiles"]
    for w, wn in enumerate(("warp 0 (sub 0, seg-sum)", "warp 9 (sub 1, copy-out)")):
        tiles = t[:, w, 5].clamp_min(1)
        print(wn, {n: round(float((t[:, w, i] / tiles).mean()), 1) for i, n in enumerate(names[:5])}, "tiles/CTA", float(tiles.mean()),
              file=sys.stderr)
    _msg.DBG_TIMING = None
print(json.dumps({"edges": g.e, "nodes": g.n + g.m, "mode": a.mode,
                  "kernels": {k: {"ms": v[0] / v[2], "GBps": v[1] / max(v[0], 1e-9) / 1e6} for k, v in
                              sorted(agg.items(), key=lambda kv: -kv[1][0]) if k.startswith(("msg", "graph.tr"))}}))
