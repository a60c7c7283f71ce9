"""N>1 host logic on CPU: world_size-2 gloo run of the flat-gradient all-reduce and the parameter broadcast used by
TrainStep (the Morton-range decomposition itself is covered by test_domain_cpu.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from se3gnn_b200.dist import allreduce_mean_, broadcast_params_, flatten_grads
    torch.manual_seed(rank)                       # replicas start DIFFERENT; the broadcast makes them rank 0's
    lin = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    broadcast_params_(lin.parameters())
    torch.manual_seed(0)
    ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    assert all(torch.equal(a, b) for a, b in zip(lin.parameters(), ref.parameters()))
    flat = flatten_grads(lin.parameters())
    x = torch.full((4, 5), float(rank + 1))
    lin(x).sum().backward()                       # autograd accumulates in place into the flat views
    assert flat.abs().sum() > 0
    local = flat.clone()
    allreduce_mean_(flat)
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    ok = torch.allclose(flat, sum(gathered) / world, atol=1e-6)
    same_views = all(p.grad.data_ptr() >= flat.data_ptr() for p in lin.parameters())
    q.put((rank, bool(ok), bool(same_views)))
    dist.destroy_process_group()


def test_flat_grad_allreduce_gloo_world2():
    import sys
    from conftest import PKG
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    os.environ["PYTHONPATH"] = PKG + os.pathsep + os.environ.get("PYTHONPATH", "")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1] and all(r[1] and r[2] for r in res)
