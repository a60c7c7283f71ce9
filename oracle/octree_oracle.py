"""TEST INFRASTRUCTURE ONLY — CPU specification of the octree graph builder.

PARITY UNPINNED: the reference's octree / edge-generation source is NOT in the mount
(/root/reference holds only models/segnn/l1_tensor_prod.py), so this
file is a *self-authored specification* written from BASELINE.json's north_star text
("numba-jit tree build, cell-to-particle assignment, hierarchical edge/neighbour
generation").  The CUDA builder must reproduce it bit for bit (integer work); tests pin
this oracle itself with brute-force O(N^2) / O(M^2) checks.  Only tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.

Specification
-------------
1. Bounding cube.  lo[a] = min_i pos[i,a];  L = max_a(max_i pos[i,a] - lo[a]) (fp32; 1 if 0).
   scale = fl32(2^21 / L);  q[i,a] = min(floor(fl32(fl32(pos[i,a]-lo[a]) * scale)), 2^21-1).
2. Morton key (63 bit): bit b of x,y,z goes to key bits 3b+2, 3b+1, 3b.
3. Stable sort by key (ties by original index): order[r] = original index of rank r.
4. Octree, level-synchronous.  Cell 0 = root = ranks [0,N), level 0.  A cell splits iff
   count > leaf_size and level < max_depth; its children are the non-empty octants in
   increasing digit order.  Cells are numbered breadth-first (level by level, Morton order
   within a level).  leaf_of_rank[r] = id of the leaf holding rank r.
5. Graph nodes = N particles (by rank) followed by M cells (node id N+c).  Directed edges
   dst <- src, stored CSR by dst with sources ascending:
     particle r in leaf c   <- every other particle of c ; <- node N+c
     cell c                 <- its particles (leaf only) ; <- parent ; <- same-level cells whose
                               integer coordinates differ by at most 1 per axis (26-neighbourhood)
                               ; <- its children.
   The edge set is symmetric.
6. Cell moments (bottom-up): mass = sum m, pos = sum(m x)/mass, vel = sum(m v)/mass.
"""
from __future__ import annotations

import numpy as np
from numba import njit

MAX_DEPTH = 21


def quantize(pos: np.ndarray):
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    lo = pos.min(axis=0)
    ext = (pos.max(axis=0) - lo).astype(np.float32)
    L = np.float32(ext.max())
    if not (L > 0):
        L = np.float32(1.0)
    scale = np.float32(np.float32(2097152.0) / L)
    t = ((pos - lo).astype(np.float32) * scale).astype(np.float32)
    q = np.minimum(np.floor(t).astype(np.int64), (1 << 21) - 1).astype(np.uint64)
    return q, lo, scale


def _spread3(v: np.ndarray) -> np.ndarray:
    v = v.astype(np.uint64) & np.uint64(0x1FFFFF)
    v = (v | (v << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
    v = (v | (v << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
    v = (v | (v << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
    v = (v | (v << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
    v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
    return v


def morton_keys(pos: np.ndarray) -> np.ndarray:
    q, _, _ = quantize(pos)
    return (_spread3(q[:, 0]) << np.uint64(2)) | (_spread3(q[:, 1]) << np.uint64(1)) | _spread3(q[:, 2])


@njit(cache=True)
def _build_tree(keys, leaf_size, max_depth):
    n = keys.shape[0]
    cap = max(16, 2 * n + 64)
    start = np.zeros(cap, np.int32)
    count = np.zeros(cap, np.int32)
    level = np.zeros(cap, np.int32)
    parent = np.full(cap, -1, np.int32)
    first_child = np.full(cap, -1, np.int32)
    nchild = np.zeros(cap, np.int32)
    ckey = np.zeros(cap, np.uint64)
    level_ptr = np.zeros(max_depth + 2, np.int32)
    m = 1
    count[0] = n
    level_ptr[0] = 0
    level_ptr[1] = 1
    nlev = 1
    lev = 0
    while True:
        beg, end = level_ptr[lev], level_ptr[lev + 1]
        made = 0
        if lev < max_depth:
            shift = np.uint64(3 * (max_depth - lev - 1))
            for c in range(beg, end):
                if count[c] <= leaf_size:
                    continue
                s, e = start[c], start[c] + count[c]
                first_child[c] = m
                i = s
                while i < e:
                    d = (keys[i] >> shift) & np.uint64(7)
                    j = i + 1
                    while j < e and ((keys[j] >> shift) & np.uint64(7)) == d:
                        j += 1
                    if m >= cap:
                        raise ValueError("cell capacity")
                    start[m] = i
                    count[m] = j - i
                    level[m] = lev + 1
                    parent[m] = c
                    ckey[m] = keys[i] >> shift
                    m += 1
                    nchild[c] += 1
                    made += 1
                    i = j
        if made == 0:
            break
        lev += 1
        level_ptr[lev + 1] = m
        nlev += 1
    return (start[:m].copy(), count[:m].copy(), level[:m].copy(), parent[:m].copy(), first_child[:m].copy(),
            nchild[:m].copy(), ckey[:m].copy(), level_ptr[:nlev + 1].copy())


@njit(cache=True)
def _compact3(k):
    # inverse of spread3 for one coordinate (bits 0,3,6,...)
    x = np.uint64(0)
    for b in range(21):
        x |= ((k >> np.uint64(3 * b)) & np.uint64(1)) << np.uint64(b)
    return x


@njit(cache=True)
def _spread1(v):
    x = np.uint64(0)
    for b in range(21):
        x |= ((v >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b)
    return x


@njit(cache=True)
def _find(ckey, beg, end, key):
    lo, hi = beg, end
    while lo < hi:
        mid = (lo + hi) // 2
        if ckey[mid] < key:
            lo = mid + 1
        else:
            hi = mid
    if lo < end and ckey[lo] == key:
        return lo
    return -1


@njit(cache=True)
def _cell_neighbours(c, level, ckey, level_ptr, out):
    """ids of same-level cells in the 26-neighbourhood of c, ascending; returns how many."""
    lev = level[c]
    if lev == 0:
        return 0
    beg, end = level_ptr[lev], level_ptr[lev + 1]
    k = ckey[c]
    cx, cy, cz = np.int64(_compact3(k >> np.uint64(2))), np.int64(_compact3(k >> np.uint64(1))), np.int64(_compact3(k))
    lim = np.int64(1) << lev
    cnt = 0
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                if dx == 0 and dy == 0 and dz == 0:
                    continue
                x, y, z = cx + dx, cy + dy, cz + dz
                if x < 0 or y < 0 or z < 0 or x >= lim or y >= lim or z >= lim:
                    continue
                nk = (_spread1(np.uint64(x)) << np.uint64(2)) | (_spread1(np.uint64(y)) << np.uint64(1)) | _spread1(np.uint64(z))
                j = _find(ckey, beg, end, nk)
                if j >= 0:
                    out[cnt] = j
                    cnt += 1
    # insertion sort (<= 26 entries)
    for a in range(1, cnt):
        v = out[a]
        b = a - 1
        while b >= 0 and out[b] > v:
            out[b + 1] = out[b]
            b -= 1
        out[b + 1] = v
    return cnt


@njit(cache=True)
def _build_edges(n, start, count, level, parent, first_child, nchild, ckey, level_ptr, leaf_of_rank):
    m = start.shape[0]
    nn = n + m
    rowptr = np.zeros(nn + 1, np.int64)
    nb = np.zeros(26, np.int32)
    # degrees
    for r in range(n):
        rowptr[r + 1] = count[leaf_of_rank[r]]  # (count-1) particles + the leaf cell
    for c in range(m):
        d = 0
        if first_child[c] < 0:
            d += count[c]
        if parent[c] >= 0:
            d += 1
        d += _cell_neighbours(c, level, ckey, level_ptr, nb)
        d += nchild[c]
        rowptr[n + c + 1] = d
    for i in range(nn):
        rowptr[i + 1] += rowptr[i]
    col = np.empty(rowptr[nn], np.int32)
    for r in range(n):
        c = leaf_of_rank[r]
        p = rowptr[r]
        for j in range(start[c], start[c] + count[c]):
            if j != r:
                col[p] = j
                p += 1
        col[p] = n + c
    for c in range(m):
        p = rowptr[n + c]
        if first_child[c] < 0:
            for j in range(start[c], start[c] + count[c]):
                col[p] = j
                p += 1
        if parent[c] >= 0:
            col[p] = n + parent[c]
            p += 1
        k = _cell_neighbours(c, level, ckey, level_ptr, nb)
        for a in range(k):
            col[p] = n + nb[a]
            p += 1
        for a in range(nchild[c]):
            col[p] = n + first_child[c] + a
            p += 1
    return rowptr, col


def build_graph(pos: np.ndarray, leaf_size: int = 32, max_depth: int = MAX_DEPTH):
    """Full CPU specification.  Returns a dict of numpy arrays (see module docstring)."""
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    n = pos.shape[0]
    keys = morton_keys(pos)
    order = np.argsort(keys, kind="stable").astype(np.int32)
    skeys = keys[order]
    start, count, level, parent, first_child, nchild, ckey, level_ptr = _build_tree(skeys, leaf_size, max_depth)
    m = start.shape[0]
    leaf_of_rank = np.empty(n, np.int32)
    leaves = np.nonzero(first_child < 0)[0]
    for c in leaves:
        leaf_of_rank[start[c]:start[c] + count[c]] = c
    cell_of_particle = np.empty(n, np.int32)
    cell_of_particle[order] = leaf_of_rank
    rowptr, col = _build_edges(n, start, count, level, parent, first_child, nchild, ckey, level_ptr, leaf_of_rank)
    dst = np.repeat(np.arange(n + m, dtype=np.int32), np.diff(rowptr).astype(np.int64))
    return dict(n=n, m=m, keys=skeys, order=order, cell_start=start, cell_count=count, cell_level=level,
                cell_parent=parent, cell_first_child=first_child, cell_nchild=nchild, cell_key=ckey,
                level_ptr=level_ptr, leaf_of_rank=leaf_of_rank, cell_of_particle=cell_of_particle,
                rowptr=rowptr, col=col, dst=dst)


def cell_moments(g, pos, vel, mass):
    """Per-cell (mass, com, mean velocity) in fp64 (tolerance-checked, not bit-exact)."""
    o = g["order"]
    p, v, w = pos[o].astype(np.float64), vel[o].astype(np.float64), mass[o].astype(np.float64)
    cw = np.concatenate([[0.0], np.cumsum(w)])
    cp = np.concatenate([np.zeros((1, 3)), np.cumsum(p * w[:, None], 0)])
    cv = np.concatenate([np.zeros((1, 3)), np.cumsum(v * w[:, None], 0)])
    s, e = g["cell_start"].astype(np.int64), (g["cell_start"] + g["cell_count"]).astype(np.int64)
    mm = cw[e] - cw[s]
    return mm, (cp[e] - cp[s]) / mm[:, None], (cv[e] - cv[s]) / mm[:, None]
