"""BASELINE configs[2] as a measured case (first version: fp32 SIMT tensor products reading their gathered inputs in
place, gate kernels; the aggregation is still torch's index_add): octree graph build + SH(2) attributes + 4-layer SEGNN l_max = 2 forward/backward + Adam on one GPU.
One JSON line; CUDA events on the current stream, warm-up first.  The model keeps the per-edge tensor-product outputs
for autograd (about 4.8 KB per edge over 4 layers).

    python tools/bench_segnn_l2.py [--particles 100000] [--steps 5]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scalable-e3-gnn_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=100_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--hidden", default="23x0e+7x1o+4x2e")
    a = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from models.segnn.segnn_l2 import SEGNNL2
    from se3gnn_b200 import capi
    from se3gnn_b200.octree import build_octree_graph, sh2_attributes
    rng = np.random.default_rng(1)
    n = a.particles
    u = rng.random(n)
    r = np.minimum(1.0 / np.sqrt(u ** (-2.0 / 3.0) - 1.0), 10.0)
    d = rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pos = torch.from_numpy((r[:, None] * d).astype(np.float32)).cuda()
    vel = torch.randn(n, 3, device="cuda")
    torch.manual_seed(0)
    model = SEGNNL2(a.hidden, 4).cuda()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
    target = torch.randn(n, 3, device="cuda")

    def step():
        g = build_octree_graph(pos, vel, leaf_size=32)
        out = model.forward_graph(g, sh2_attributes(g))
        loss = (out[:n] - target).square().mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return g, loss

    for _ in range(a.warmup):
        g, loss = step()
    torch.cuda.synchronize()
    n0 = capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        g, loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    nl = (capi.launch_count() - n0) / a.steps
    # one more step with CUDA events around every library call: where the time goes (outside the timed region)
    capi.profile_begin()
    step()
    agg = {}
    for tag, t, _, _ in capi.profile_end():
        k = agg.setdefault(tag, [0.0, 0])
        k[0] += t
        k[1] += 1
    kernels = {k: {"ms_per_step": round(v[0], 3), "calls": v[1]} for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}
    mode = os.environ.get("SE3_L2_MSG", "table")
    print(json.dumps({
        "workload": f"SEGNN l_max=2, 4 layers, hidden {a.hidden}, {n} particles (plummer), fp32, octree leaf size 32 "
                    + ("[BASELINE configs[2], fp32 contraction]" if n == 1_000_000 else "[BASELINE configs[2] at a different size]"),
        "ms_per_step": round(ms, 3), "particles_per_s": n / ms * 1e3, "edges": int(g.e), "cells": int(g.m),
        "gpu_launches_per_step": nl, "loss": float(loss.detach()), "message1": mode,
        "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2),
        "kernels": kernels,
        "note": "tensor products csrc/o3tp.cu (fp32 SIMT, gathered inputs read in place), message 1 "
                + ("by linearity (node tables + per-edge coupling, csrc/o3msg.cu)" if mode == "table" else "as a tensor product on the concatenated row")
                + ", gates csrc/gate.cu; aggregation (index_add) and residual are torch ops",
    }), flush=True)


if __name__ == "__main__":
    main()
