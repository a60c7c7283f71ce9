"""Octree graph builder (host side).  Spec: ``oracle/octree_oracle.py``; kernels:
``csrc/octree.cu``.  The reference's graph-builder API is not observable (its source is not in
the mount), so this PyG-style API is builder-defined:

    g = build_octree_graph(pos, vel, mass, leaf_size=32)
    g.edge_index            # [2, E] int32 (src, dst), sorted by (dst, src), node ids = Morton ranks then cells
    g.order                 # rank -> original particle index
    g.cell_of_particle      # leaf cell id per particle (original order)

PyTorch provides device memory and the stream only; every array is filled by CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import torch

from . import capi

MAX_DEPTH = 21


@dataclass
class OctreeGraph:
    n: int
    m: int
    e: int
    nlevels: int
    leaf_size: int
    keys: torch.Tensor            # [n] int64 (bit pattern of the uint64 Morton keys), sorted
    order: torch.Tensor           # [n] int32
    cell_start: torch.Tensor
    cell_count: torch.Tensor
    cell_level: torch.Tensor
    cell_parent: torch.Tensor
    cell_first_child: torch.Tensor
    cell_nchild: torch.Tensor
    cell_key: torch.Tensor
    level_ptr: torch.Tensor       # [23] int32
    leaf_of_rank: torch.Tensor
    cell_of_particle: torch.Tensor
    bbox: torch.Tensor            # lo xyz, scale
    rowptr: torch.Tensor          # [n+m+1] int64
    col: torch.Tensor             # [e] int32 source node
    dst: torch.Tensor             # [e] int32 target node
    node_pos: Optional[torch.Tensor] = None
    node_vel: Optional[torch.Tensor] = None
    node_mass: Optional[torch.Tensor] = None
    edge_attr: Optional[torch.Tensor] = None    # [e,4]
    edge_extra: Optional[torch.Tensor] = None   # [e,2]
    node_attr: Optional[torch.Tensor] = None    # [n+m,4]
    x_in: Optional[torch.Tensor] = None         # [n+m,8] = 2x1o + 2x0e

    @property
    def num_nodes(self) -> int:
        return self.n + self.m

    @property
    def edge_index(self) -> torch.Tensor:
        return torch.stack([self.col, self.dst])


def _need_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"se3gnn_b200.octree: {name} must be a CUDA tensor (there is no CPU fallback)")


def build_octree_graph(pos: torch.Tensor, vel: Optional[torch.Tensor] = None, mass: Optional[torch.Tensor] = None,
                       leaf_size: int = 32, cell_cap: Optional[int] = None, features: bool = True) -> OctreeGraph:
    _need_cuda(pos, "pos")
    lib = capi.lib()
    pos = pos.to(torch.float32).contiguous()
    n = int(pos.shape[0])
    if pos.dim() != 2 or pos.shape[1] != 3 or n < 1:
        raise ValueError("pos must be [n,3] with n >= 1")
    dev = pos.device
    cap = int(cell_cap) if cell_cap is not None else 2 * n + 64
    i32 = dict(device=dev, dtype=torch.int32)
    keys = torch.empty(n, device=dev, dtype=torch.int64)
    order = torch.empty(n, **i32)
    cells = torch.empty((6, cap), **i32)
    cell_key = torch.empty(cap, device=dev, dtype=torch.int64)
    level_ptr = torch.empty(MAX_DEPTH + 2, **i32)
    leaf_of_rank = torch.empty(n, **i32)
    cell_of_particle = torch.empty(n, **i32)
    bbox = torch.empty(4, device=dev, dtype=torch.float32)
    wb = C.c_size_t()
    capi.check(lib.se3_octree_work_bytes(n, cap, C.byref(wb)))
    work = torch.empty(wb.value, device=dev, dtype=torch.uint8)
    t = capi.Octree()
    t.n, t.leaf_size, t.max_depth, t.cell_cap = n, int(leaf_size), MAX_DEPTH, cap
    t.keys, t.order = keys.data_ptr(), order.data_ptr()
    (t.cell_start, t.cell_count, t.cell_level, t.cell_parent, t.cell_first_child, t.cell_nchild) = (
        cells[i].data_ptr() for i in range(6))
    t.cell_key, t.level_ptr = cell_key.data_ptr(), level_ptr.data_ptr()
    t.leaf_of_rank, t.cell_of_particle, t.bbox = leaf_of_rank.data_ptr(), cell_of_particle.data_ptr(), bbox.data_ptr()
    t.work, t.work_bytes = work.data_ptr(), wb.value
    st = capi.current_stream_ptr()
    m_out, nlev = C.c_int64(), C.c_int32()
    # algorithmic bytes: keys 12+12 B/particle, 8 sort passes x (8 hist + 12 read + 12 write) B,
    # split/assign ~ (8 read + 4 write) B per particle per level touched (charged once here)
    with capi.mark("octree.build", n * (24.0 + 8 * 32.0 + 12.0)):
        capi.check(lib.se3_octree_build(pos.data_ptr(), C.byref(t), C.byref(m_out), C.byref(nlev), st), "se3_octree_build")
    m = int(m_out.value)
    nn = n + m
    nbr = torch.empty((m, 26), **i32)
    deg = torch.empty(nn, **i32)
    rowptr = torch.empty(nn + 1, device=dev, dtype=torch.int64)
    scan_work = torch.empty(nn // 1024 + 2, device=dev, dtype=torch.int64)
    e_out = C.c_int64()
    with capi.mark("graph.degrees", m * (8.0 + 104.0 + 4.0) + n * 8.0 + nn * 12.0):
        capi.check(lib.se3_graph_degrees(C.byref(t), m, nbr.data_ptr(), deg.data_ptr(), rowptr.data_ptr(),
                                         scan_work.data_ptr(), C.byref(e_out), st), "se3_graph_degrees")
    e = int(e_out.value)
    col = torch.empty(e, **i32)
    dst = torch.empty(e, **i32)
    with capi.mark("graph.emit", e * 8.0 + m * 104.0 + nn * 8.0):
        capi.check(lib.se3_graph_emit(C.byref(t), m, nbr.data_ptr(), rowptr.data_ptr(), col.data_ptr(), dst.data_ptr(), st),
                   "se3_graph_emit")
    g = OctreeGraph(n=n, m=m, e=e, nlevels=int(nlev.value), leaf_size=int(leaf_size), keys=keys, order=order,
                    cell_start=cells[0, :m], cell_count=cells[1, :m], cell_level=cells[2, :m], cell_parent=cells[3, :m],
                    cell_first_child=cells[4, :m], cell_nchild=cells[5, :m], cell_key=cell_key[:m], level_ptr=level_ptr,
                    leaf_of_rank=leaf_of_rank, cell_of_particle=cell_of_particle, bbox=bbox, rowptr=rowptr, col=col, dst=dst)
    if features:
        if vel is None:
            vel = torch.zeros_like(pos)
        if mass is None:
            mass = torch.full((n,), 1.0 / n, device=dev, dtype=torch.float32)
        _need_cuda(vel, "vel")
        _need_cuda(mass, "mass")
        vel = vel.to(torch.float32).contiguous()
        mass = mass.to(torch.float32).contiguous()
        f32 = dict(device=dev, dtype=torch.float32)
        g.node_pos = torch.empty((nn, 3), **f32)
        g.node_vel = torch.empty((nn, 3), **f32)
        g.node_mass = torch.empty(nn, **f32)
        with capi.mark("graph.node_data", n * 60.0 + m * 56.0):
            capi.check(lib.se3_node_data(C.byref(t), m, g.nlevels, pos.data_ptr(), vel.data_ptr(), mass.data_ptr(),
                                         g.node_pos.data_ptr(), g.node_vel.data_ptr(), g.node_mass.data_ptr(), st),
                       "se3_node_data")
        g.edge_attr = torch.empty((e, 4), **f32)
        g.edge_extra = torch.empty((e, 2), **f32)
        g.node_attr = torch.empty((nn, 4), **f32)
        g.x_in = torch.empty((nn, 8), **f32)
        with capi.mark("graph.edge_geometry", e * (8.0 + 12.0 + 24.0 + 16.0) + nn * (28.0 + 48.0)):
            capi.check(lib.se3_edge_geometry(n, m, e, rowptr.data_ptr(), col.data_ptr(), dst.data_ptr(),
                                             g.node_pos.data_ptr(), g.node_vel.data_ptr(), g.node_mass.data_ptr(),
                                             C.c_float(float(n)), g.edge_attr.data_ptr(), g.edge_extra.data_ptr(),
                                             g.node_attr.data_ptr(), g.x_in.data_ptr(), st), "se3_edge_geometry")
    return g


def sh2_attributes(g: OctreeGraph):
    """SH(2) edge / node attributes of a graph built with features: (edge_attr9 [E,9], node_attr9 [N+M,9]) for the l <= 2
    tensor product (`se3gnn_b200.o3tp`); the first four columns equal `g.edge_attr` / `g.node_attr`."""
    if getattr(g, "node_pos", None) is None:
        raise capi.Se3Error("sh2_attributes needs a graph built with features=True")
    dev = g.node_pos.device
    nn = g.n + g.m
    ea = torch.empty((g.e, 9), device=dev, dtype=torch.float32)
    na = torch.empty((nn, 9), device=dev, dtype=torch.float32)
    with capi.mark("graph.edge_geometry_l2", g.e * (8.0 + 24.0 + 36.0 + 36.0) + nn * (12.0 + 36.0)):
        capi.check(capi.lib().se3_edge_geometry_l2(g.n, g.m, g.e, g.rowptr.data_ptr(), g.col.data_ptr(), g.dst.data_ptr(),
                                                   g.node_pos.data_ptr(), g.node_vel.data_ptr(), ea.data_ptr(),
                                                   na.data_ptr(), capi.current_stream_ptr()), "se3_edge_geometry_l2")
    return ea, na
