"""Pin the self-authored octree/graph specification (oracle/octree_oracle.py) with brute-force
checks at small N.  (The reference's builder is not in the mount: parity vs the reference is
unpinned; these tests make the spec itself trustworthy.)"""
import numpy as np
import pytest

from oracle import octree_oracle as T


def _clouds():
    rng = np.random.default_rng(0)
    yield "uniform", rng.random((700, 3)).astype(np.float32)
    r = rng.random(900) ** 3
    d = rng.standard_normal((900, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    yield "clustered", (r[:, None] * d).astype(np.float32)
    p = rng.random((300, 3)).astype(np.float32)
    p[:80] = p[0]  # 80 coincident points -> a max-depth leaf above leaf_size
    yield "duplicates", p
    yield "tiny", rng.random((5, 3)).astype(np.float32)
    yield "single", np.zeros((1, 3), np.float32)
    yield "line", np.stack([np.linspace(0, 1, 200), np.zeros(200), np.zeros(200)], 1).astype(np.float32)


@pytest.mark.parametrize("name,pos", list(_clouds()), ids=[n for n, _ in _clouds()])
def test_tree_invariants(name, pos):
    g = T.build_graph(pos, leaf_size=8)
    n, m = g["n"], g["m"]
    keys = g["keys"]
    assert np.all(keys[:-1] <= keys[1:])
    # stable: ties keep original order
    o = g["order"]
    ties = keys[:-1] == keys[1:]
    assert np.all(o[:-1][ties] < o[1:][ties])
    assert sorted(o.tolist()) == list(range(n))
    st, ct, lv, pa, fc, nc = (g[k] for k in ("cell_start", "cell_count", "cell_level", "cell_parent",
                                              "cell_first_child", "cell_nchild"))
    leaves = np.nonzero(fc < 0)[0]
    # leaves partition the ranks
    cover = np.zeros(n, int)
    for c in leaves:
        cover[st[c]:st[c] + ct[c]] += 1
        assert ct[c] <= 8 or lv[c] == T.MAX_DEPTH
        assert np.all(g["leaf_of_rank"][st[c]:st[c] + ct[c]] == c)
    assert np.all(cover == 1)
    for c in range(m):
        if fc[c] >= 0:
            assert ct[c] > 8 and lv[c] < T.MAX_DEPTH
            ch = np.arange(fc[c], fc[c] + nc[c])
            assert np.all(pa[ch] == c) and np.all(lv[ch] == lv[c] + 1)
            assert ct[ch].sum() == ct[c] and st[ch[0]] == st[c]
            assert np.all(np.diff(g["cell_key"][ch].astype(np.int64)) > 0)
        # every particle of the cell carries the cell's key prefix
        sh = np.uint64(3 * (T.MAX_DEPTH - lv[c]))
        assert np.all((keys[st[c]:st[c] + ct[c]] >> sh) == g["cell_key"][c])
    # BFS numbering: levels are contiguous and Morton-sorted inside a level
    lp = g["level_ptr"]
    for l in range(len(lp) - 1):
        ids = np.arange(lp[l], lp[l + 1])
        assert np.all(lv[ids] == l)
        assert np.all(np.diff(g["cell_key"][ids].astype(np.int64)) > 0)
    assert np.all(g["cell_of_particle"][o] == g["leaf_of_rank"])


def _coords(key, lev):
    x = y = z = 0
    for b in range(lev):
        x |= ((int(key) >> (3 * b + 2)) & 1) << b
        y |= ((int(key) >> (3 * b + 1)) & 1) << b
        z |= ((int(key) >> (3 * b)) & 1) << b
    return x, y, z


@pytest.mark.parametrize("name,pos", list(_clouds()), ids=[n for n, _ in _clouds()])
def test_edges_bruteforce(name, pos):
    g = T.build_graph(pos, leaf_size=8)
    n, m = g["n"], g["m"]
    want = set()
    lor = g["leaf_of_rank"]
    for r in range(n):
        c = lor[r]
        for j in range(g["cell_start"][c], g["cell_start"][c] + g["cell_count"][c]):
            if j != r:
                want.add((r, j))
        want.add((r, n + c))
        want.add((n + c, r))
    crd = [_coords(g["cell_key"][c], g["cell_level"][c]) for c in range(m)]
    for c in range(m):
        if g["cell_parent"][c] >= 0:
            want.add((n + c, n + g["cell_parent"][c]))
            want.add((n + g["cell_parent"][c], n + c))
        for d in range(m):
            if d != c and g["cell_level"][c] == g["cell_level"][d] and \
                    max(abs(a - b) for a, b in zip(crd[c], crd[d])) <= 1:
                want.add((n + c, n + d))
    got = list(zip(g["dst"].tolist(), g["col"].tolist()))
    assert len(got) == len(set(got)) == len(want)
    assert set(got) == want
    assert got == sorted(got)                                   # canonical (dst, src) order
    assert all((s, d) in want for d, s in got)                  # symmetric
    rp = g["rowptr"]
    assert rp[0] == 0 and rp[-1] == len(got) and np.all(np.diff(rp) >= 0)


def test_quantisation_corner_cases():
    pos = np.array([[0, 0, 0], [1, 1, 1], [0.5, 0.25, 1.0]], np.float32)
    q, lo, scale = T.quantize(pos)
    assert q.max() == (1 << 21) - 1 and q.min() == 0
    k = T.morton_keys(pos)
    assert k[0] == 0 and k[1] == (1 << 63) - 1


def test_cell_moments():
    rng = np.random.default_rng(3)
    pos = rng.random((400, 3)).astype(np.float32)
    vel = rng.standard_normal((400, 3)).astype(np.float32)
    mass = rng.random(400).astype(np.float32)
    g = T.build_graph(pos, leaf_size=8)
    mm, com, cv = T.cell_moments(g, pos, vel, mass)
    assert mm[0] == pytest.approx(mass.astype(np.float64).sum())
    np.testing.assert_allclose(com[0], (pos * mass[:, None]).sum(0) / mass.sum(), rtol=1e-5)
    c = g["m"] - 1
    idx = g["order"][g["cell_start"][c]:g["cell_start"][c] + g["cell_count"][c]]
    np.testing.assert_allclose(cv[c], (vel[idx] * mass[idx, None]).sum(0) / mass[idx].sum(), rtol=1e-5)
