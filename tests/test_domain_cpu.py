"""Morton-range domain decomposition (se3gnn_b200.domain) on CPU: world-2 and world-3 gloo runs of the CPU oracle models
(l_max = 1, and the l_max = 2 model of BASELINE configs[2]) on the oracle's octree graph, partitioned and halo-exchanged
by the PRODUCT's host logic, against the 1-rank result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _cloud(n, seed=3):
    rng = np.random.default_rng(seed)
    d = rng.standard_normal((n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = np.minimum(1.0 / np.sqrt(np.maximum(rng.random(n), 1e-12) ** (-2.0 / 3.0) - 1.0), 10.0)
    pos = (r[:, None] * d).astype(np.float32)
    vel = rng.standard_normal((n, 3)).astype(np.float32)
    mass = np.full(n, 1.0 / n, np.float32)
    return pos, vel, mass


def _global(n, leaf):
    from oracle import octree_oracle as T
    from oracle.segnn_oracle import graph_features
    pos, vel, mass = _cloud(n)
    g = T.build_graph(pos, leaf_size=leaf)
    f = graph_features(g, pos, vel, mass)
    t = lambda a, dt=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dt)
    return g, dict(x_in=t(f["x_in"]), node_attr=t(f["node_attr"]), edge_attr=t(f["edge_attr"]), edge_extra=t(f["edge_extra"]),
                   dst=t(g["dst"], torch.int64), src=t(g["col"], torch.int64),
                   cell_start=t(g["cell_start"], torch.int64), leaf_of_rank=t(g["leaf_of_rank"], torch.int64))


def _model(l2=False):
    torch.manual_seed(0)
    if l2:
        from oracle.segnn_l2_oracle import SEGNNL2Oracle
        return SEGNNL2Oracle(hidden="5x0e+2x1o+1x2e", num_layers=2)
    from oracle.segnn_oracle import SEGNNOracle
    return SEGNNOracle(hidden="6x0e+3x1o", num_layers=2).double()


def _global_l2(n, leaf):
    """The same graph with SH(2) attributes (BASELINE configs[2] model)."""
    from oracle.segnn_l2_oracle import sh2_features
    from oracle.segnn_oracle import graph_features
    g, G = _global(n, leaf)
    pos, vel, mass = _cloud(n)
    ea, na = sh2_features(graph_features(g, pos, vel, mass), g)
    G["edge_attr"], G["node_attr"] = torch.from_numpy(ea), torch.from_numpy(na)
    return g, G


def _worker(rank, world, port, n, leaf, q, l2=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from se3gnn_b200 import domain
    g, G = _global_l2(n, leaf) if l2 else _global(n, leaf)
    lg = domain.local_graph(rank, world, g["n"], G["cell_start"], G["leaf_of_rank"], G["dst"], G["src"])
    domain.exchange_halo_lists(lg)
    model = _model(l2)
    out = model(G["x_in"][lg.own_ids], G["node_attr"][lg.own_ids], G["edge_attr"][lg.edge_ids], G["edge_extra"][lg.edge_ids],
                lg.dst, lg.src, halo=lambda x: domain.halo_exchange(x, lg))
    loss = out[:lg.n_part].square().sum() / (3.0 * n)
    loss.backward()
    flat = torch.cat([p.grad.flatten() for p in model.parameters()])
    dist.all_reduce(flat)
    ld = loss.detach().clone()
    dist.all_reduce(ld)
    q.put((rank, lg.part_lo, lg.n_part, lg.n_own, lg.n_halo, lg.e, out[:lg.n_part].detach().numpy(), flat.numpy(), float(ld),
           sum(lg.send_counts), sum(lg.recv_counts)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,leaf,l2", [(2, 1500, 16, False), (3, 900, 8, False), (2, 600, 8, True)])
def test_decomposed_matches_single_rank(world, n, leaf, l2):
    from conftest import PKG, ROOT
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    os.environ["PYTHONPATH"] = os.pathsep.join([PKG, ROOT, os.path.join(ROOT, "tests"), os.environ.get("PYTHONPATH", "")])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, leaf, q, l2)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
    # single-rank reference
    g, G = _global_l2(n, leaf) if l2 else _global(n, leaf)
    model = _model(l2)
    out = model(G["x_in"], G["node_attr"], G["edge_attr"], G["edge_extra"], G["dst"], G["src"])
    loss = out[:n].square().sum() / (3.0 * n)
    loss.backward()
    flat = torch.cat([p.grad.flatten() for p in model.parameters()]).numpy()
    ref = out[:n].detach().numpy()
    assert sum(r[2] for r in res) == n                                   # every particle owned exactly once
    assert sum(r[3] for r in res) == g["n"] + g["m"]                     # every node owned exactly once
    assert sum(r[5] for r in res) == len(g["dst"])                       # every edge kept exactly once
    assert sum(r[9] for r in res) == sum(r[10] for r in res) > 0         # halo rows sent == received, and there are some
    lo = 0
    for r in res:
        assert r[1] == lo
        np.testing.assert_allclose(r[6], ref[lo:lo + r[2]], rtol=1e-9, atol=1e-12)
        lo += r[2]
        np.testing.assert_allclose(r[7], flat, rtol=1e-8, atol=1e-12)
        assert abs(r[8] - float(loss)) < 1e-12


def test_slab_bounds_are_leaf_aligned():
    from se3gnn_b200 import domain
    g, G = _global(700, 8)
    for world in (2, 4, 8):
        b = domain.slab_bounds(g["n"], world, G["leaf_of_rank"], G["cell_start"])
        assert b[0] == 0 and b[-1] == g["n"] and bool((b[1:] >= b[:-1]).all())
        leaf_starts = set(int(G["cell_start"][c]) for c in np.nonzero(g["cell_first_child"] < 0)[0])
        assert all(int(x) in leaf_starts or int(x) == g["n"] for x in b[1:-1])
        own = domain.node_owner(b, g["n"], G["cell_start"])
        # a leaf and all its particles have one owner
        assert bool((own[:g["n"]] == own[g["n"] + G["leaf_of_rank"]]).all())


def test_gather_rows_and_runs():
    from se3gnn_b200 import domain
    torch.manual_seed(0)
    for w in (2, 4, 8):
        t = torch.randn(777, w)
        ids = torch.randint(0, 777, (3000,))
        assert torch.equal(domain.gather_rows(t, ids), t[ids])
    ids = torch.tensor([3, 4, 5, 9, 10, 20])
    assert domain._runs(ids) == ([3, 9, 20], [6, 11, 21])
    assert domain._runs(torch.tensor([7])) == ([7], [8])
    assert domain._runs(torch.empty(0, dtype=torch.int64)) == ([], [])


def test_row_range_extraction_equals_mask_extraction():
    from se3gnn_b200 import domain
    g, G = _global(1200, 16)
    rp = torch.from_numpy(g["rowptr"].astype(np.int64))
    for world in (2, 5):
        for r in range(world):
            a = domain.local_graph(r, world, g["n"], G["cell_start"], G["leaf_of_rank"], G["dst"], G["src"])
            b = domain.local_graph(r, world, g["n"], G["cell_start"], G["leaf_of_rank"], G["dst"], G["src"], rowptr=rp)
            assert torch.equal(a.edge_ids, b.edge_ids) and torch.equal(a.dst, b.dst) and torch.equal(a.src, b.src)
            assert torch.equal(a.halo_ids, b.halo_ids) and a.recv_counts == b.recv_counts
            assert torch.equal(domain.take_edges(b, G["edge_attr"]), domain.take_edges(a, G["edge_attr"]))
