#!/usr/bin/env python
"""Multi-GPU check of the Morton-range domain decomposition over NCCL (run under torchrun, one rank per GPU):
the decomposed step must reproduce rank 0's single-GPU loss and weight gradients on the same cloud."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "scalable-e3-gnn_b200")):
    sys.path.insert(0, p)
import numpy as np, torch, torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from models.segnn.segnn import SEGNN
from se3gnn_b200.pipeline import TrainStep, synthetic_cloud
from se3gnn_b200 import domain
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 4
data = [torch.from_numpy(a).cuda() for a in synthetic_cloud(n, "plummer", seed=7)]
torch.manual_seed(0); m1 = SEGNN(num_layers=layers).cuda()
ts = TrainStep(m1, decompose=True)
loss = float(ts.step_device_dd(*data)); torch.cuda.synchronize()
lg = ts.last_local
info = torch.tensor([lg.n_part, lg.n_own, lg.n_halo, lg.e, sum(lg.send_counts)], device="cuda", dtype=torch.int64)
allinfo = [torch.zeros_like(info) for _ in range(world)]
dist.all_gather(allinfo, info)
if rank == 0:
    torch.manual_seed(0); m0 = SEGNN(num_layers=layers).cuda()
    t0 = TrainStep(m0)
    ref = float(t0.step_device(*data)); torch.cuda.synchronize()
    err = (ts.flat_grad - t0.flat_grad).abs().max().item() / t0.flat_grad.abs().max().item()
    print(f"world {world} n {n}: loss dd {loss:.8f} single {ref:.8f} rel {abs(loss-ref)/abs(ref):.2e}; grad rel err {err:.2e}")
    for r, t in enumerate(allinfo):
        p, o, h, e, s = t.tolist()
        print(f"  rank {r}: particles {p} owned nodes {o} halo {h} edges {e} rows sent/layer {s} "
              f"({domain.halo_bytes(lg, 64, layers) / 1e6:.2f} MB/step on rank 0)" if r == 0 else
              f"  rank {r}: particles {p} owned nodes {o} halo {h} edges {e} rows sent/layer {s}")
    # weight gradients are fp32 sums over ~10^6 edges with heavy cancellation: a different partition of the rows changes
    # the reduction order (per-CTA TMEM accumulators, atomics); run-to-run noise of ONE configuration is ~2e-5 of max|g|,
    # between decompositions ~1e-4 (2 ranks) .. ~1e-3 (8 ranks, 3.6M edges).  The halo logic itself is checked exactly
    # (fp64, 1e-9) by tests/test_domain_cpu.py.
    assert abs(loss - ref) <= 2e-5 * abs(ref) and err < 5e-3
    print("check_dd ok")
dist.destroy_process_group()
