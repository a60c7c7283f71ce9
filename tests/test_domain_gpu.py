"""Morton-range domain decomposition on the GPU path: two ranks (two processes sharing cuda:0, gloo rendezvous with
host-staged halo exchange — NCCL refuses two ranks on one device; the multi-GPU NCCL run is tools/check_dd.py)
run TrainStep.step_device_dd on one cloud and must reproduce the single-rank loss and weight gradients."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _setup(n, layers):
    from models.segnn.segnn import SEGNN
    from se3gnn_b200.pipeline import TrainStep, synthetic_cloud
    torch.manual_seed(0)
    model = SEGNN(num_layers=layers).cuda()
    data = [torch.from_numpy(a).cuda() for a in synthetic_cloud(n, "plummer", seed=5)]
    return model, data, TrainStep


def _worker(rank, world, port, n, layers, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    from se3gnn_b200 import capi
    model, data, TrainStep = _setup(n, layers)
    ts = TrainStep(model, decompose=True)
    l0 = capi.launch_count()
    loss = ts.step_device_dd(*data)
    torch.cuda.synchronize()
    lg = ts.last_local
    q.put((rank, float(loss), ts.flat_grad.cpu().numpy(), lg.n_part, lg.n_halo, lg.e, capi.launch_count() - l0))
    dist.destroy_process_group()


@pytest.mark.parametrize("n,layers", [(20000, 2)])
def test_two_rank_decomposition_matches_single_rank(n, layers):
    from conftest import PKG, ROOT
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    os.environ["PYTHONPATH"] = os.pathsep.join([PKG, ROOT, os.path.join(ROOT, "tests"), os.environ.get("PYTHONPATH", "")])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, layers, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    model, data, TrainStep = _setup(n, layers)
    ts = TrainStep(model)
    loss = float(ts.step_device(*data))
    torch.cuda.synchronize()
    ref = ts.flat_grad.cpu().numpy()
    g = ts.last_graph
    assert res[0][3] + res[1][3] == n and res[0][5] + res[1][5] == g.e
    assert res[0][4] > 0 and res[1][4] > 0 and all(r[6] > 20 for r in res)   # halos exist, CUDA kernels ran
    scale = np.abs(ref).max()
    for r in res:
        assert abs(r[1] - loss) <= 2e-5 * abs(loss)
        assert np.abs(r[2] - ref).max() <= 1e-4 * scale


@pytest.mark.parametrize("n,world", [(20000, 2), (30011, 3), (50000, 8), (300, 8)])
def test_cuda_local_graph_equals_torch_twin(n, world):
    """csrc/domain.cu (ownership, local CSR, halo lists grouped by owner) against the torch implementation that the CPU
    gloo tests run: every array identical, for every rank of the decomposition (no collectives needed: the global graph
    is replicated)."""
    from se3gnn_b200 import domain
    from se3gnn_b200.octree import build_octree_graph
    from se3gnn_b200.pipeline import synthetic_cloud
    pos, vel, mass, _ = (torch.from_numpy(a).cuda() for a in synthetic_cloud(n, "plummer", seed=3))
    g = build_octree_graph(pos, vel, mass)
    tot_own = tot_e = 0
    for rank in range(world):
        ref = domain.local_graph(rank, world, g.n, g.cell_start, g.leaf_of_rank, g.dst, g.col)
        lg = domain.local_graph_cuda(rank, world, g)
        domain.finish_halo(lg)          # world > 1 without a process group: only the local part
        eq = lambda a, b: torch.testing.assert_close(a.long().cpu(), b.long().cpu(), rtol=0, atol=0)
        assert (lg.n_part, lg.n_own, lg.e, lg.n_halo, lg.part_lo) == (ref.n_part, ref.n_own, ref.e, ref.n_halo, ref.part_lo)
        eq(lg.own_ids, ref.own_ids)
        eq(lg.halo_ids, ref.halo_ids)
        eq(lg.dst, ref.dst)
        eq(lg.src, ref.src)
        assert lg.recv_counts == ref.recv_counts
        torch.testing.assert_close(lg.edge_attr, domain.take_edges(ref, g.edge_attr), rtol=0, atol=0)
        torch.testing.assert_close(lg.edge_extra, domain.take_edges(ref, g.edge_extra), rtol=0, atol=0)
        rp = torch.searchsorted(lg.dst.long().contiguous(), torch.arange(lg.n_own + 1, device="cuda"))
        eq(lg.rowptr, rp)
        tot_own += lg.n_own
        tot_e += lg.e
    assert tot_own == g.n + g.m and tot_e == g.e


def _worker_l2(rank, world, port, n, layers, q):
    """l_max = 2 model on a decomposed graph: message 1 by linearity with halo rows (n_all > n_dst), gate + aggregation
    over the owned destinations, halo all-to-all per layer (gloo, host staged)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    from models.segnn.segnn_l2 import SEGNNL2
    from se3gnn_b200 import capi, domain
    from se3gnn_b200.octree import build_octree_graph, sh2_attributes
    from se3gnn_b200.pipeline import synthetic_cloud
    torch.manual_seed(0)
    model = SEGNNL2("8x0e+3x1o+2x2e", layers).cuda()
    pos, vel, mass, target = [torch.from_numpy(a).cuda() for a in synthetic_cloud(n, "plummer", seed=5)]
    g = build_octree_graph(pos, vel, mass, leaf_size=16)
    ea, na = sh2_attributes(g)
    lg = domain.local_graph(rank, world, g.n, g.cell_start, g.leaf_of_rank, g.dst, g.col, rowptr=g.rowptr)
    domain.exchange_halo_lists(lg)
    own = lg.own_ids.long()
    l0 = capi.launch_count()
    out = model(g.x_in.index_select(0, own), na.index_select(0, own), ea.index_select(0, lg.edge_ids),
                g.edge_extra.index_select(0, lg.edge_ids), lg.dst, lg.src, halo=lambda x: domain.halo_exchange(x, lg))
    tgt = target.index_select(0, g.order[lg.part_lo:lg.part_lo + lg.n_part].long())
    loss = (out[:lg.n_part] - tgt).square().sum() / (3.0 * n)
    loss.backward()
    grad = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    buf = torch.cat([grad, loss.detach().reshape(1)]).cpu()
    dist.all_reduce(buf)
    torch.cuda.synchronize()
    q.put((rank, float(buf[-1]), buf[:-1].numpy(), lg.n_part, lg.n_halo, lg.e, capi.launch_count() - l0))
    dist.destroy_process_group()


def test_two_rank_decomposition_l2_model_matches_single_rank():
    from conftest import PKG, ROOT
    from models.segnn.segnn_l2 import SEGNNL2
    from se3gnn_b200.octree import build_octree_graph, sh2_attributes
    from se3gnn_b200.pipeline import synthetic_cloud
    n, layers = 6000, 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    os.environ["PYTHONPATH"] = os.pathsep.join([PKG, ROOT, os.path.join(ROOT, "tests"), os.environ.get("PYTHONPATH", "")])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_l2, args=(r, 2, port, n, layers, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    torch.manual_seed(0)
    model = SEGNNL2("8x0e+3x1o+2x2e", layers).cuda()
    pos, vel, mass, target = [torch.from_numpy(a).cuda() for a in synthetic_cloud(n, "plummer", seed=5)]
    g = build_octree_graph(pos, vel, mass, leaf_size=16)
    out = model.forward_graph(g, sh2_attributes(g))
    loss = (out[:n] - target.index_select(0, g.order.long())).square().mean()
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).cpu().numpy()
    assert res[0][3] + res[1][3] == n and res[0][5] + res[1][5] == g.e
    assert res[0][4] > 0 and res[1][4] > 0 and all(r[6] > 20 for r in res)   # halos exist, CUDA kernels ran
    scale = np.abs(ref).max()
    lossf = float(loss.detach())
    for r in res:
        assert abs(r[1] - lossf) <= 2e-5 * abs(lossf)
        assert np.abs(r[2] - ref).max() <= 5e-4 * scale
