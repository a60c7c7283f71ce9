#!/usr/bin/env python
"""Octree graph construction only (BASELINE configs[4]): clustered (Plummer / NFW-like) and uniform clouds,
sweep of sizes on one GPU; edges/s and algorithmic GB/s against the measured HBM peak.  One JSON line per case.
usage: bench_octree.py [sizes comma-separated] [kinds comma-separated] [reps]
Several GPUs (configs[4] "at 1/8 GPUs"): the build does not shard below one cloud (the decomposed SEGNN step replicates it,
DESIGN 6), so under torchrun every rank builds its OWN cloud of the given size (replicas only, no collective on the data
path) and rank 0 prints the aggregate: points and edges of all ranks over the slowest rank's time.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_octree.py 1e7,1e8 plummer"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scalable-e3-gnn_b200")):
    sys.path.insert(0, p)
import torch
from se3gnn_b200 import capi
from se3gnn_b200.octree import build_octree_graph

sizes = [int(float(x)) for x in (sys.argv[1] if len(sys.argv) > 1 else "1e6,1e7").split(",")]
kinds = (sys.argv[2] if len(sys.argv) > 2 else "plummer,uniform").split(",")
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
peak = 6550.1
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)


def cloud(n, kind):
    g = torch.Generator(device=dev); g.manual_seed(1 + rank)
    if kind == "uniform":
        return torch.rand((n, 3), device=dev, generator=g)
    d = torch.randn((n, 3), device=dev, generator=g)
    d /= d.norm(dim=1, keepdim=True)
    u = torch.rand(n, device=dev, generator=g).clamp_min(1e-12)
    if kind == "plummer":
        r = (1.0 / torch.sqrt(u ** (-2.0 / 3.0) - 1.0)).clamp_max(10.0)
    else:  # "nfw"-like cusp: r ~ u^2 (steeper central concentration than Plummer), truncated at 1
        r = u * u
    return (r[:, None] * d).contiguous()


for kind in kinds:
    for n in sizes:
        pos = cloud(n, kind)
        torch.cuda.synchronize()
        best, g = None, None
        for _ in range(reps + 1):
            del g
            capi.profile_begin()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g = build_octree_graph(pos, leaf_size=32, features=False)
            e1.record(); torch.cuda.synchronize()
            prof = capi.profile_end()
            ms = e0.elapsed_time(e1)
            if best is None or ms < best[0]:
                best = (ms, prof)
        ms, prof = best
        nb = sum(p[2] for p in prof)
        edges, cells = g.e, g.m
        if world > 1:   # replicas: totals over the ranks, time of the slowest rank
            t = torch.tensor([float(edges), float(cells), float(nb)], device=dev, dtype=torch.float64)
            dist.all_reduce(t)
            tm = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            edges, cells, nb, ms = int(t[0].item()), int(t[1].item()), float(t[2].item()), float(tm.item())
        line = {"workload": f"octree graph build only, {n} {kind} points" + (f" per GPU x {world} GPUs (independent clouds)" if world > 1 else "") + ", leaf 32",
                "n_gpus": world, "particles": n * world, "cells": cells, "edges": edges,
                "levels": g.nlevels, "ms": ms, "edges_per_s": edges / (ms * 1e-3), "particles_per_s": n * world / (ms * 1e-3),
                "algorithmic_GBps": nb / (ms * 1e-3) / 1e9, "hbm_peak_GBps": peak * world, "frac": nb / (ms * 1e-3) / 1e9 / (peak * world),
                "stages_ms": {t_: round(m_, 4) for t_, m_, _, _ in prof}}
        if rank == 0:
            print(json.dumps(line), flush=True)
        del g, pos
        torch.cuda.empty_cache()

if world > 1:
    dist.destroy_process_group()
