"""Morton-range domain decomposition of ONE particle cloud over the ranks of a box (north star).

Scheme ("replicated tree, partitioned graph"):
  * every rank holds the whole cloud (the bench all-gathers the slabs; 12 B/particle over NVLink) and builds the same
    octree — integer work, bit-identical on every rank — so ownership, halo sets and the canonical edge order need
    no negotiation;
  * particles are split by Morton rank into `world` contiguous, equal-count slabs whose boundaries are snapped down to
    a leaf start, so a leaf and all of its particles have one owner; a cell is owned by the owner of its first
    particle;
  * a rank keeps the edges whose *destination* it owns (CSR rows of its nodes), in the global (dst, src) order;
  * sources it does not own are its halo (cells across the boundary: 26-neighbours, parents, children).  Before every
    message layer the halo rows of the node features are fetched from their owners with one all-to-all-v (NCCL over
    NVLink / NVSwitch: every peer at full bandwidth, so a single grouped exchange, no ring ordering); backward sends
    the halo gradients the opposite way and adds them into the owners' rows;
  * weight gradients are summed over ranks with one all-reduce of the flat gradient buffer.

Everything here is index bookkeeping on torch tensors of whatever device the graph lives on (CUDA in the product,
CPU in the world-2 gloo tests, which run the same code on the oracle's numpy graph).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def slab_bounds(n: int, world: int, leaf_of_rank: torch.Tensor, cell_start: torch.Tensor) -> torch.Tensor:
    """[world+1] int64 particle-rank boundaries: equal counts (remainder to the first ranks), each interior boundary
    snapped DOWN to the first rank of the leaf that contains it.  Monotone; slabs may be empty for tiny clouds."""
    base, rem = divmod(n, world)
    raw = [0]
    for r in range(world):
        raw.append(raw[-1] + base + (1 if r < rem else 0))
    b = torch.tensor(raw, dtype=torch.int64, device=leaf_of_rank.device)
    inner = b[1:-1]
    if inner.numel():
        ok = inner < n
        leaf = leaf_of_rank[inner.clamp(max=max(n - 1, 0))].long()
        snapped = torch.where(ok, cell_start[leaf].long(), inner)
        b[1:-1] = snapped
    return b


@dataclass
class LocalGraph:
    """What one rank keeps of the global octree graph.  Local node ids: owned particles (global rank order), owned
    cells (global cell order), then halo nodes (ascending global id)."""
    rank: int
    world: int
    n_global: int                  # particles of the whole cloud
    nn_global: int
    n_part: int                    # owned particles
    n_own: int                     # owned nodes (particles + cells)
    n_halo: int
    part_lo: int                   # global rank of the first owned particle
    own_ids: torch.Tensor          # [n_own]  global node ids
    halo_ids: torch.Tensor         # [n_halo] global node ids, ascending
    dst: torch.Tensor              # [e_loc] int32 local (owned) destination, ascending
    src: torch.Tensor              # [e_loc] int32 local source (owned or halo)
    edge_ids: torch.Tensor         # [e_loc] int64 positions in the global edge arrays
    recv_counts: List[int] = field(default_factory=list)   # halo rows received from each rank (halo order)
    send_counts: List[int] = field(default_factory=list)   # rows sent to each rank
    send_idx: Optional[torch.Tensor] = None                # [sum(send_counts)] int64 local owned ids, grouped by rank
    edge_runs: Optional[list] = None                       # [(a, b)] ranges of the global edge arrays (CSR rows of owned nodes)
    rowptr: Optional[torch.Tensor] = None                  # [n_own + 1] int64 CSR row pointers of dst (CUDA path)
    edge_attr: Optional[torch.Tensor] = None               # [e_loc, 4] the rank's rows of the per-edge arrays (CUDA path)
    edge_extra: Optional[torch.Tensor] = None              # [e_loc, 2]
    pos: Optional[torch.Tensor] = None                     # [nn] int32 global node -> local id of owned nodes, -1 otherwise
    counts: Optional[torch.Tensor] = None                  # device int64 [4 + world] (CUDA path, before finish_halo)

    @property
    def e(self) -> int:
        return int(self.dst.numel())


def node_owner(bounds: torch.Tensor, n: int, cell_start: torch.Tensor) -> torch.Tensor:
    """[n+m] int64 owner rank of every global node."""
    dev = cell_start.device
    first = torch.cat([torch.arange(n, device=dev, dtype=torch.int64), cell_start.long()])
    # owner = last slab whose lower bound <= first particle (empty slabs never own anything)
    own = torch.searchsorted(bounds[1:].contiguous(), first, right=True)
    return own.clamp(max=bounds.numel() - 2)


def _runs(ids: torch.Tensor):
    """Maximal runs of consecutive integers in a sorted 1-D tensor -> (starts, ends) python lists (ends exclusive)."""
    if ids.numel() == 0:
        return [], []
    brk = torch.nonzero(ids[1:] != ids[:-1] + 1).flatten() + 1
    starts = torch.cat([ids[:1], ids[brk]])
    ends = torch.cat([ids[brk - 1], ids[-1:]]) + 1
    return starts.tolist(), ends.tolist()


def local_graph(rank: int, world: int, n: int, cell_start: torch.Tensor, leaf_of_rank: torch.Tensor,
                g_dst: torch.Tensor, g_src: torch.Tensor, bounds: Optional[torch.Tensor] = None,
                rowptr: Optional[torch.Tensor] = None) -> LocalGraph:
    """Rank `rank`'s part of the global graph (no communication: the global graph is replicated).

    With `rowptr` (CSR by destination, as the builder returns it) the owned edges are taken as the CSR row ranges of
    the owned nodes — one range for the particle slab and one per octree level — instead of a mask over all E global
    edges; `LocalGraph.edge_runs` then lets callers slice per-edge arrays with a few contiguous copies."""
    dev = g_dst.device
    m = int(cell_start.numel())
    nn = n + m
    if bounds is None:
        bounds = slab_bounds(n, world, leaf_of_rank, cell_start)
    owner = node_owner(bounds, n, cell_start)
    mine = owner == rank
    own_ids = torch.nonzero(mine).flatten()
    n_own = int(own_ids.numel())
    n_part = int((own_ids < n).sum())
    runs = None
    if rowptr is not None and n_own > 0:
        ns, ne = _runs(own_ids)                                   # node ranges: particle slab + one per level
        idx = torch.tensor(ns + ne, device=dev, dtype=torch.int64)
        rp = rowptr.index_select(0, idx).tolist()
        runs = [(int(a), int(b)) for a, b in zip(rp[:len(ns)], rp[len(ns):]) if b > a]
        e_ids = torch.cat([torch.arange(a, b, device=dev, dtype=torch.int64) for a, b in runs]) if runs else \
            torch.empty(0, device=dev, dtype=torch.int64)
        sg = torch.cat([g_src[a:b] for a, b in runs]).long() if runs else e_ids
        dg = torch.cat([g_dst[a:b] for a, b in runs]).long() if runs else e_ids
    else:
        e_ids = torch.nonzero(mine.index_select(0, g_dst)).flatten()      # int32 indices: no E-sized int64 copy
        sg = g_src.index_select(0, e_ids).long()
        dg = g_dst.index_select(0, e_ids).long()
    halo_ids = torch.unique(sg[~mine[sg]])          # sorted ascending
    n_halo = int(halo_ids.numel())
    g2l = torch.full((nn,), -1, device=dev, dtype=torch.int64)
    g2l[own_ids] = torch.arange(n_own, device=dev, dtype=torch.int64)
    g2l[halo_ids] = n_own + torch.arange(n_halo, device=dev, dtype=torch.int64)
    howner = owner[halo_ids]
    recv_counts = torch.bincount(howner, minlength=world).tolist() if n_halo else [0] * world
    # halo_ids ascending is NOT grouped by owner in general (cells of several levels interleave owners): regroup
    if n_halo:
        perm = torch.argsort(howner, stable=True)
        halo_ids = halo_ids[perm]
        g2l[halo_ids] = n_own + torch.arange(n_halo, device=dev, dtype=torch.int64)
    lo = int(bounds[rank])
    lg = LocalGraph(rank=rank, world=world, n_global=n, nn_global=nn, n_part=n_part, n_own=n_own, n_halo=n_halo,
                    part_lo=lo, own_ids=own_ids, halo_ids=halo_ids, dst=g2l[dg].to(torch.int32),
                    src=g2l[sg].to(torch.int32), edge_ids=e_ids, recv_counts=[int(c) for c in recv_counts])
    lg.edge_runs = runs
    return lg


def local_graph_cuda(rank: int, world: int, g, bounds: Optional[torch.Tensor] = None) -> LocalGraph:
    """``local_graph`` for an ``OctreeGraph`` on the GPU (csrc/domain.cu): three kernels' worth of passes, ONE small D2H
    read (the sizes of the local arrays); the halo counts stay on the device for ``finish_halo``.  Same result as the
    torch version above (which remains the host-logic twin the CPU gloo tests run)."""
    import ctypes as C
    from . import capi
    lib = capi.lib()
    dev = g.dst.device
    n, m = g.n, g.m
    nn = n + m
    if bounds is None:
        bounds = slab_bounds(n, world, g.leaf_of_rank, g.cell_start)
    bounds = bounds.to(torch.int64).contiguous()
    st = capi.current_stream_ptr()
    wb = C.c_size_t()
    capi.check(lib.se3_domain_work_bytes(nn, C.byref(wb)))
    work = torch.empty(wb.value, device=dev, dtype=torch.uint8)
    pos = torch.empty(nn, device=dev, dtype=torch.int32)
    lrp = torch.empty(nn + 1, device=dev, dtype=torch.int64)
    counts = torch.empty(4 + world, device=dev, dtype=torch.int64)
    cs = g.cell_start.contiguous()
    with capi.mark("domain.mark", 8.0 * 3 * nn):
        capi.check(lib.se3_domain_mark(n, m, rank, world, bounds.data_ptr(), capi.ptr(cs), g.rowptr.data_ptr(), pos.data_ptr(),
                                       lrp.data_ptr(), counts.data_ptr(), work.data_ptr(), wb.value, st), "se3_domain_mark")
    n_part, n_own, e_loc, part_lo = (int(v) for v in torch.cat([counts[:3], bounds[rank:rank + 1]]).tolist())   # the one blocking read
    i32 = dict(device=dev, dtype=torch.int32)
    own_ids = torch.empty(n_own, **i32)
    dst_l, src_l = torch.empty(e_loc, **i32), torch.empty(e_loc, **i32)
    ea = torch.empty((e_loc, 4), device=dev, dtype=torch.float32) if g.edge_attr is not None else None
    ex = torch.empty((e_loc, 2), device=dev, dtype=torch.float32) if g.edge_extra is not None else None
    halo_ids = torch.empty(nn, **i32)
    hpos = torch.empty(nn, **i32)
    scratch = torch.empty(2 * nn, **i32)
    with capi.mark("domain.edges", 4.0 * (e_loc * (3 + 6 * 2) + 4 * nn)):
        capi.check(lib.se3_domain_edges(n, m, world, bounds.data_ptr(), capi.ptr(cs), pos.data_ptr(), lrp.data_ptr(),
                                        g.rowptr.data_ptr(), g.col.data_ptr(), capi.ptr(g.edge_attr), capi.ptr(g.edge_extra),
                                        n_own, e_loc, capi.ptr(own_ids), capi.ptr(dst_l), capi.ptr(src_l), capi.ptr(ea),
                                        capi.ptr(ex), halo_ids.data_ptr(), hpos.data_ptr(), scratch.data_ptr(),
                                        counts.data_ptr(), work.data_ptr(), wb.value, st), "se3_domain_edges")
    lg = LocalGraph(rank=rank, world=world, n_global=n, nn_global=nn, n_part=n_part, n_own=n_own, n_halo=-1,
                    part_lo=part_lo, own_ids=own_ids, halo_ids=halo_ids, dst=dst_l, src=src_l, edge_ids=None,
                    recv_counts=[])
    lg.rowptr, lg.edge_attr, lg.edge_extra, lg.pos, lg.counts = lrp[:n_own + 1], ea, ex, pos, counts
    lg.bounds = bounds
    return lg


def finish_halo(lg: LocalGraph, group=None) -> LocalGraph:
    """Second half of the CUDA path: per-owner halo counts -> every owner learns what it must send.  One all-gather of
    the count vectors, ONE blocking read (n_halo, receive and send counts together), one all-to-all-v of the ids."""
    dev = lg.own_ids.device
    world = lg.world
    rc_dev = lg.counts[4:4 + world]
    live = world > 1 and dist.is_available() and dist.is_initialized()
    if live:
        if dist.get_backend(group) == "nccl":
            mat = torch.empty((world, world), device=dev, dtype=torch.int64)
            dist.all_gather_into_tensor(mat, rc_dev.contiguous(), group=group)
        else:   # gloo (single-GPU emulation in the tests): staged through the host
            parts = [torch.empty(world, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(parts, rc_dev.cpu(), group=group)
            mat = torch.stack(parts).to(dev)
        packed = torch.cat([lg.counts[3:4], rc_dev, mat[:, lg.rank]])
    else:
        packed = torch.cat([lg.counts[3:4], rc_dev, rc_dev])
    vals = [int(v) for v in packed.tolist()]
    lg.n_halo = vals[0]
    lg.recv_counts = vals[1:1 + world]
    lg.send_counts = vals[1 + world:1 + 2 * world]
    lg.halo_ids = lg.halo_ids[:lg.n_halo]
    if live:
        want = torch.empty(sum(lg.send_counts), dtype=torch.int32, device=dev)
        _a2a(want, lg.halo_ids, lg.send_counts, lg.recv_counts, group)
        lg.send_idx = lg.pos.index_select(0, want.long()).long()
    else:
        lg.send_idx = torch.empty(0, dtype=torch.int64, device=dev)
    return lg


def gather_rows(t: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """t[ids] for a contiguous [R, w] float32 tensor.  Rows of 8 / 16 bytes are gathered as ONE 1-D index_select of
    8- / 16-byte elements (float64 / complex128 views): torch's row gather of narrow 2-D tensors was measured at
    ~0.5 ms for 1.8M rows, the 1-D path runs at memory speed."""
    if t.dim() == 2 and t.dtype == torch.float32 and t.is_contiguous() and t.shape[1] in (2, 4):
        wide = torch.float64 if t.shape[1] == 2 else torch.complex128
        return t.view(wide).reshape(-1).index_select(0, ids).view(torch.float32).reshape(-1, t.shape[1])
    return t.index_select(0, ids)


def take_edges(lg: LocalGraph, t: torch.Tensor) -> torch.Tensor:
    """Rows of a per-edge array of the global graph that belong to this rank, in local edge order."""
    runs = getattr(lg, "edge_runs", None)
    if runs is None:
        return gather_rows(t, lg.edge_ids)
    if not runs:
        return t[:0]
    return torch.cat([t[a:b] for a, b in runs])


def _a2a(out: torch.Tensor, inp: torch.Tensor, out_split: Sequence[int], in_split: Sequence[int], group=None):
    """all-to-all-v of rows.  NCCL: on the device.  gloo (CPU tests, single-GPU emulation): staged through the host."""
    if dist.get_backend(group) == "nccl" or not inp.is_cuda:
        dist.all_to_all_single(out, inp.contiguous(), output_split_sizes=list(out_split), input_split_sizes=list(in_split),
                               group=group)
        return out
    o = torch.empty(out.shape, dtype=out.dtype)
    dist.all_to_all_single(o, inp.contiguous().cpu(), output_split_sizes=list(out_split), input_split_sizes=list(in_split),
                           group=group)
    out.copy_(o)
    return out


def exchange_halo_lists(lg: LocalGraph, group=None) -> LocalGraph:
    """Tell every owner which of its nodes this rank needs; fills send_counts / send_idx.  One small all-to-all of
    counts and one of global ids, once per graph build."""
    dev = lg.own_ids.device
    world = lg.world
    rc = torch.tensor(lg.recv_counts, dtype=torch.int64, device=dev)
    sc = torch.empty(world, dtype=torch.int64, device=dev)
    _a2a(sc, rc, [1] * world, [1] * world, group)
    lg.send_counts = [int(c) for c in sc.tolist()]
    want = torch.empty(sum(lg.send_counts), dtype=torch.int64, device=dev)
    _a2a(want, lg.halo_ids, lg.send_counts, lg.recv_counts, group)
    g2l = torch.full((lg.nn_global,), -1, device=dev, dtype=torch.int64)
    g2l[lg.own_ids] = torch.arange(lg.n_own, device=dev, dtype=torch.int64)
    lg.send_idx = g2l[want]
    if lg.send_idx.numel() and int(lg.send_idx.min()) < 0:
        raise RuntimeError("domain decomposition: a peer asked for a node this rank does not own")
    return lg


class _HaloFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_own: torch.Tensor, lg: LocalGraph, group):
        ctx.lg, ctx.group = lg, group
        d = x_own.shape[1]
        send = x_own.index_select(0, lg.send_idx)
        recv = torch.empty((lg.n_halo, d), device=x_own.device, dtype=x_own.dtype)
        _a2a(recv, send, lg.recv_counts, lg.send_counts, group)
        return torch.cat([x_own, recv], 0)

    @staticmethod
    def backward(ctx, g_ext: torch.Tensor):
        lg = ctx.lg
        g_own = g_ext[:lg.n_own].clone()
        g_halo = g_ext[lg.n_own:].contiguous()
        back = torch.empty((int(lg.send_idx.numel()), g_ext.shape[1]), device=g_ext.device, dtype=g_ext.dtype)
        _a2a(back, g_halo, lg.send_counts, lg.recv_counts, ctx.group)
        g_own.index_add_(0, lg.send_idx, back)
        return g_own, None, None


def halo_exchange(x_own: torch.Tensor, lg: LocalGraph, group=None) -> torch.Tensor:
    """[n_own, D] -> [n_own + n_halo, D]: owned rows followed by the halo rows fetched from their owners
    (differentiable: backward returns the halo gradients to the owners and adds them)."""
    if lg.world == 1:
        return x_own
    return _HaloFn.apply(x_own, lg, group)


def halo_bytes(lg: LocalGraph, d: int, layers: int) -> int:
    """NVLink bytes this rank sends per training step (forward rows + backward gradients), fp32."""
    return 4 * d * layers * (sum(lg.send_counts) + sum(lg.recv_counts))
