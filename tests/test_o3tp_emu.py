"""CPU emulation of the l <= 2 tensor-product CUDA tile programs (`tests/emu/o3tp_emu.cpp` compiles the kernels' own
source, `csrc/o3tp_body.inl` + `csrc/o3tp_tables.h`, with the block's threads run sequentially) against the oracle.
Checks the planning, the table walk and all shared-memory indexing without a GPU; the GPU parity test proper is
`tests/test_o3tp_gpu.py`."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import lmax2_oracle as l2

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "o3tp_emu.cpp")
DEPS = [SRC] + [os.path.join(HERE, "..", "scalable-e3-gnn_b200", "csrc", f) for f in ("o3tp_body.inl", "o3tp_tables.h", "o3tp_cg_gen.inl")]
LIB = os.path.join(HERE, "emu", "_build", "libo3tp_emu.so")


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in DEPS):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", SRC, "-o", LIB], check=True)
    return C.CDLL(LIB)


def _flat(irreps, pair=False):
    v = []
    for mul, l, p in irreps:
        v += [l, p] if pair else [mul, l, p]
    return (C.c_int * len(v))(*v)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Segs:
    """C arrays describing in1 as row segments (+ gradient destinations) for the emulation entry points."""

    def __init__(self, parts, grads=None, modes=None):
        # parts: [(array [n, ld] float32, idx int32 array or None, width)]
        n = len(parts)
        self.keep = parts, grads
        self.base = (C.POINTER(C.c_float) * 4)(*[_fp(a) for a, _, _ in parts])
        self.idx = (C.POINTER(C.c_int) * 4)(*[i.ctypes.data_as(C.POINTER(C.c_int)) if i is not None else None for _, i, _ in parts])
        self.width = (C.c_int * 4)(*[w for _, _, w in parts])
        self.ld = (C.c_int * 4)(*[a.shape[1] for a, _, _ in parts])
        self.n = n
        if grads is not None:
            self.gbase = (C.POINTER(C.c_float) * 4)(*[_fp(g) if g is not None else None for g in grads])
            self.gmode = (C.c_int * 4)(*modes)

    def fwd(self):
        return self.n, self.base, self.idx, self.width, self.ld

    def bwd(self):
        return self.n, self.base, self.idx, self.width, self.ld, self.gbase, self.gmode


CASES = {
    "sh1": ([(8, 0, 1), (4, 1, -1)], 1, [(6, 0, 1), (5, 1, -1)]),
    "balanced2": ([(23, 0, 1), (7, 1, -1), (4, 2, 1)], 2, [(23, 0, 1), (7, 1, -1), (4, 2, 1)]),
    "message2": ([(23, 0, 1), (7, 1, -1), (4, 2, 1), (23, 0, 1), (7, 1, -1), (4, 2, 1), (2, 0, 1)], 2,
                 [(34, 0, 1), (7, 1, -1), (4, 2, 1)]),
    "mixed_parity": ([(3, 0, 1), (2, 1, -1), (2, 2, 1), (1, 1, 1), (1, 2, -1), (2, 0, -1)], 2,
                     [(3, 0, 1), (2, 1, -1), (1, 2, 1), (2, 1, 1), (1, 2, -1), (5, 0, -1)]),
    "dead_output": ([(4, 0, 1)], 1, [(3, 0, 1), (2, 1, 1), (2, 1, -1)]),       # 1e has no path: columns must be 0
    "scalar_attr": ([(5, 0, 1), (3, 2, 1)], 0, [(4, 0, 1), (2, 2, 1)]),
    "wide": ([(64, 0, 1), (32, 1, -1), (16, 2, 1)], 2, [(64, 0, 1), (32, 1, -1), (16, 2, 1)]),
}


def _oracle(in1, in2, out, x1, y, w, g):
    ws, o = [], 0
    for shp in l2.weight_shapes(in1, in2, out):
        n = shp[0] * shp[1]
        ws.append(torch.from_numpy(w[o:o + n].astype(np.float64).reshape(shp)).requires_grad_())
        o += n
    assert o == len(w)
    x1t = torch.from_numpy(x1.astype(np.float64)).requires_grad_()
    yt = torch.from_numpy(y.astype(np.float64)).requires_grad_()
    res = l2.forward(x1t, yt, ws, in1, in2, out)
    (res * torch.from_numpy(g.astype(np.float64))).sum().backward()
    gw = np.concatenate([t.grad.numpy().reshape(-1) for t in ws])
    return res.detach().numpy(), x1t.grad.numpy(), yt.grad.numpy(), gw


def _close(got, want, tol=2e-5):
    scale = max(1.0, float(np.abs(want).max()))
    assert np.abs(got - want).max() <= tol * scale, float(np.abs(got - want).max() / scale)


def test_couplings_match_oracle(emu):
    for a in range(3):
        for b in range(3):
            for c in range(3):
                buf = (C.c_double * 125)()
                rc = emu.emu_coupling(a, b, c, buf)
                want = l2.cg(a, b, c)
                if not want.any():
                    assert rc == -1
                    continue
                assert rc == 0
                got = np.frombuffer(buf, dtype=np.float64)[:want.size].reshape(want.shape)
                np.testing.assert_allclose(got, want, atol=1e-12)


@pytest.mark.parametrize("name", list(CASES))
def test_plan_matches_oracle(emu, name):
    in1, lmax, out = CASES[name]
    in2 = l2.sh_irreps(lmax)
    dims = (C.c_int * 4)()
    arr = [(C.c_int * 128)() for _ in range(4)]
    pa = (C.c_float * 128)()
    nw = emu.emu_plan(len(in1), _flat(in1), len(in2), _flat(in2, True), len(out), _flat(out), dims, *arr, pa)
    ps = l2.paths(in1, in2, out)
    assert dims[3] == len(ps)
    assert [(arr[1][k], arr[2][k], arr[0][k]) for k in range(len(ps))] == ps
    shapes = l2.weight_shapes(in1, in2, out)
    assert nw == sum(a * b for a, b in shapes)
    offs = np.cumsum([0] + [a * b for a, b in shapes])[:-1]
    assert [arr[3][k] for k in range(len(ps))] == list(offs)
    a = l2.norm_factors(in1, in2, out)
    np.testing.assert_allclose([pa[k] for k in range(len(ps))], [a[p[2]] for p in ps], rtol=1e-6)


@pytest.mark.parametrize("name,rows,TEF,nblocks", [
    ("sh1", 37, 32, 2), ("balanced2", 145, 64, 1), ("balanced2", 33, 32, 3), ("message2", 121, 64, 2),
    ("mixed_parity", 19, 32, 1), ("dead_output", 9, 64, 1), ("scalar_attr", 16, 32, 1), ("message2", 1, 32, 4),
    ("wide", 70, 64, 1),
])
def test_emulated_kernels_match_oracle(emu, name, rows, TEF, nblocks):
    in1, lmax, out = CASES[name]
    in2 = l2.sh_irreps(lmax)
    rng = np.random.default_rng(sum(map(ord, name)))          # stable across processes (str hashes are salted)
    d1 = sum(m * (2 * l + 1) for m, l, _ in in1)
    d2 = sum(2 * l + 1 for _, l, _ in in2)
    do = sum(m * (2 * l + 1) for m, l, _ in out)
    x1 = rng.standard_normal((rows, d1)).astype(np.float32)
    y = rng.standard_normal((rows, d2)).astype(np.float32)
    nw = sum(a * b for a, b in l2.weight_shapes(in1, in2, out))
    w = rng.standard_normal(nw).astype(np.float32)
    g = rng.standard_normal((rows, do)).astype(np.float32)
    want_o, want_gx, want_gy, want_gw = _oracle(in1, in2, out, x1, y, w, g)

    spec = (len(in1), _flat(in1), len(in2), _flat(in2, True), len(out), _flat(out))
    got_o = np.full((rows, do), np.nan, np.float32)
    # forward: one warp per (output irrep, channel chunk, 32-row group); the schedule is made for 8 warps
    assert emu.emu_forward(*spec, C.c_longlong(rows), *Segs([(x1, None, d1)]).fwd(), _fp(y), _fp(w), _fp(got_o), TEF, 256,
                           nblocks) == 0
    _close(got_o, want_o)

    gx = np.full((rows, d1), np.nan, np.float32)
    gy = np.full((rows, d2), np.nan, np.float32)
    gw = np.full(nw, np.nan, np.float32)
    assert emu.emu_backward(*spec, C.c_longlong(rows), *Segs([(x1, None, d1)], [gx], [1]).bwd(), _fp(y), _fp(w), _fp(g),
                            _fp(gy), _fp(gw), 256, nblocks) == 0
    _close(gx, want_gx)
    _close(gy, want_gy)
    _close(gw, want_gw)
    # the split backward (input-gradient kernel + weight-gradient kernel), where the plan allows it
    gx3 = np.full((rows, d1), np.nan, np.float32)
    gy3 = np.full((rows, d2), np.nan, np.float32)
    gw3 = np.full(nw, np.nan, np.float32)
    rc = emu.emu_backward_split(*spec, C.c_longlong(rows), *Segs([(x1, None, d1)], [gx3], [1]).bwd(), _fp(y), _fp(w), _fp(g),
                                _fp(gy3), _fp(gw3), 256, nblocks)
    assert rc == (1 if name in ("mixed_parity", "wide") else 0)
    if rc == 0:
        _close(gx3, want_gx)
        _close(gy3, want_gy)
        _close(gw3, want_gw)
        gx4 = np.full((rows, d1), np.nan, np.float32)
        assert emu.emu_backward_split(*spec, C.c_longlong(rows), *Segs([(x1, None, d1)], [gx4], [1]).bwd(), _fp(y), _fp(w),
                                      _fp(g), None, _fp(gw3), 256, nblocks) == 0
        np.testing.assert_array_equal(gx4, gx3)
    # gin2 is optional
    gx2 = np.full((rows, d1), np.nan, np.float32)
    gw2 = np.full(nw, np.nan, np.float32)
    assert emu.emu_backward(*spec, C.c_longlong(rows), *Segs([(x1, None, d1)], [gx2], [1]).bwd(), _fp(y), _fp(w), _fp(g),
                            None, _fp(gw2), 256, nblocks) == 0
    np.testing.assert_array_equal(gx2, gx)
    np.testing.assert_array_equal(gw2, gw)


@pytest.mark.parametrize("split,modes", [(True, [3 | 16, 2 | 16, 0]), (False, [3, 2, 0]), (True, [2, 3 | 16, 0])])
def test_emulated_gathered_segments(emu, split, modes):
    """in1 = cat(x[dst], x[src], extra) read through row segments; gradients: atomic adds into per-source buffers for the
    gathered segments, a plain store for the identity segment, one segment skipped."""
    in1 = [(6, 0, 1), (3, 1, -1), (2, 2, 1), (6, 0, 1), (3, 1, -1), (2, 2, 1), (2, 0, 1)]
    out = [(9, 0, 1), (3, 1, -1), (2, 2, 1)]
    in2 = l2.sh_irreps(2)
    rng = np.random.default_rng(7)
    nn_, rows, dh = 23, 150, 6 + 9 + 10
    pad = 3 if not (modes[0] | modes[1]) & 16 else 0                     # ld > width: padded node rows (scalar adds)
    in1 = [(9, 0, 1) if m == 6 else (m, l, p) for m, l, p in in1] if not pad else in1   # 16-byte path: width 28
    dh = sum(m * (2 * l + 1) for m, l, p in in1[:3])
    x = rng.standard_normal((nn_, dh + pad)).astype(np.float32)
    extra = rng.standard_normal((rows, 2)).astype(np.float32)
    dst = np.sort(rng.integers(0, nn_, rows)).astype(np.int32)
    src = rng.integers(0, nn_, rows).astype(np.int32)
    if modes[1] & 15 == 3:
        src = np.sort(src)
    y = rng.standard_normal((rows, 9)).astype(np.float32)
    x1 = np.concatenate([x[dst, :dh], x[src, :dh], extra], 1)
    d1, do = x1.shape[1], 9 + 9 + 10
    nw = sum(a * b for a, b in l2.weight_shapes(in1, in2, out))
    w = rng.standard_normal(nw).astype(np.float32)
    g = rng.standard_normal((rows, do)).astype(np.float32)
    want_o, want_gx, _, want_gw = _oracle(in1, in2, out, x1, y, w, g)
    spec = (len(in1), _flat(in1), len(in2), _flat(in2, True), len(out), _flat(out))
    parts = [(x, dst, dh), (x, src, dh), (extra, None, 2)]
    got_o = np.full((rows, do), np.nan, np.float32)
    assert emu.emu_forward(*spec, C.c_longlong(rows), *Segs(parts).fwd(), _fp(y), _fp(w), _fp(got_o), 64, 256, 2) == 0
    _close(got_o, want_o)
    ga, gb = np.zeros_like(x), np.zeros_like(x)
    gw = np.full(nw, np.nan, np.float32)
    fn = emu.emu_backward_split if split else emu.emu_backward
    assert fn(*spec, C.c_longlong(rows), *Segs(parts, [ga, gb, None], modes).bwd(), _fp(y), _fp(w), _fp(g), None,
              _fp(gw), 256, 2) == 0
    wa, wb = np.zeros((nn_, dh)), np.zeros((nn_, dh))
    np.add.at(wa, dst, want_gx[:, :dh])
    np.add.at(wb, src, want_gx[:, dh:2 * dh])
    _close(ga[:, :dh], wa)
    _close(gb[:, :dh], wb)
    assert not ga[:, dh:].any() and not gb[:, dh:].any()
    _close(gw, want_gw)


@pytest.mark.parametrize("order", [1, 2, 3])
def test_regions_do_not_depend_on_thread_order(emu, order):
    """The emulation runs the threads of a region one after another; a missing barrier (a thread reading what another
    thread of the same region writes) would make the result depend on that order.  Re-run representative cases with the
    threads reversed, warp-interleaved and permuted."""
    emu.emu_set_order(order)
    try:
        test_emulated_kernels_match_oracle(emu, "message2", 121, 64, 2)
        test_emulated_kernels_match_oracle(emu, "mixed_parity", 19, 32, 1)
        test_emulated_gathered_segments(emu, True, [3 | 16, 2 | 16, 0])
        test_emulated_gathered_segments(emu, False, [3, 2, 0])
        test_emulated_kernels_random_irreps(emu, 3)
    finally:
        emu.emu_set_order(0)


@pytest.mark.parametrize("seed", range(12))
def test_emulated_kernels_random_irreps(emu, seed):
    """Seeded random configurations (irreps of both parities in any order, repeated irreps, 1..8 of them, in2 = SH(0..2),
    ragged row counts): forward, fused backward and - where the plan allows it - split backward against the oracle."""
    rng = np.random.default_rng(1000 + seed)

    def rand_irreps(k):
        return [(int(rng.integers(1, 10)), int(rng.integers(0, 3)), int(rng.choice([-1, 1]))) for _ in range(k)]

    lmax = int(rng.integers(0, 3))
    in2 = l2.sh_irreps(lmax)
    for _ in range(50):
        in1, out = rand_irreps(int(rng.integers(1, 9))), rand_irreps(int(rng.integers(1, 9)))
        if l2.paths(in1, in2, out):
            break
    else:
        pytest.skip("no connected configuration drawn")
    rows = int(rng.integers(1, 150))
    d1 = sum(m * (2 * l + 1) for m, l, _ in in1)
    d2 = sum(2 * l + 1 for _, l, _ in in2)
    do = sum(m * (2 * l + 1) for m, l, _ in out)
    x1 = rng.standard_normal((rows, d1)).astype(np.float32)
    y = rng.standard_normal((rows, d2)).astype(np.float32)
    nw = sum(a * b for a, b in l2.weight_shapes(in1, in2, out))
    w = rng.standard_normal(nw).astype(np.float32)
    g = rng.standard_normal((rows, do)).astype(np.float32)
    want_o, want_gx, want_gy, want_gw = _oracle(in1, in2, out, x1, y, w, g)
    spec = (len(in1), _flat(in1), len(in2), _flat(in2, True), len(out), _flat(out))
    nblocks = int(rng.integers(1, 4))
    got_o = np.full((rows, do), np.nan, np.float32)
    assert emu.emu_forward(*spec, C.c_longlong(rows), *Segs([(x1, None, d1)]).fwd(), _fp(y), _fp(w), _fp(got_o),
                           int(rng.choice([32, 64])), 256, nblocks) == 0
    _close(got_o, want_o)
    for fn in (emu.emu_backward, emu.emu_backward_split):
        gx = np.full((rows, d1), np.nan, np.float32)
        gy = np.full((rows, d2), np.nan, np.float32)
        gw = np.full(nw, np.nan, np.float32)
        rc = fn(*spec, C.c_longlong(rows), *Segs([(x1, None, d1)], [gx], [1]).bwd(), _fp(y), _fp(w), _fp(g), _fp(gy),
                _fp(gw), 256, nblocks)
        if rc == 1 and fn is emu.emu_backward_split:
            continue                                  # plan not eligible for the split (many output irreps)
        assert rc == 0
        _close(gx, want_gx)
        _close(gy, want_gy)
        _close(gw, want_gw)


def test_emulated_kernels_match_frozen_vectors(emu):
    """The emulated kernels against the frozen convention vectors (`tests/golden/o3tp_*.npz`)."""
    import glob
    import json
    files = sorted(f for f in glob.glob(os.path.join(HERE, "golden", "o3tp_*.npz")) if not f.endswith("couplings.npz"))
    assert len(files) == 3
    for f in files:
        r = np.load(f)
        meta = json.loads(bytes(r["meta"]).decode())
        in1, out = [tuple(t) for t in meta["in1"]], [tuple(t) for t in meta["out"]]
        in2 = l2.sh_irreps(meta["lmax"])
        x1, y, w, g = (np.ascontiguousarray(r[k]) for k in ("x1", "y", "w", "g"))
        rows, d1, d2 = x1.shape[0], x1.shape[1], y.shape[1]
        spec = (len(in1), _flat(in1), len(in2), _flat(in2, True), len(out), _flat(out))
        got = np.full(r["out_f64"].shape, np.nan, np.float32)
        assert emu.emu_forward(*spec, C.c_longlong(rows), *Segs([(x1, None, d1)]).fwd(), _fp(y), _fp(w), _fp(got), 32, 256, 1) == 0
        _close(got, r["out_f64"])
        gx, gy, gw = np.full((rows, d1), np.nan, np.float32), np.full((rows, d2), np.nan, np.float32), np.full(len(w), np.nan, np.float32)
        assert emu.emu_backward(*spec, C.c_longlong(rows), *Segs([(x1, None, d1)], [gx], [1]).bwd(), _fp(y), _fp(w), _fp(g),
                                _fp(gy), _fp(gw), 256, 1) == 0
        _close(gx, r["gx_f64"])
        _close(gy, r["gy_f64"])
        _close(gw, r["gw_f64"])


@pytest.mark.parametrize("seed", range(8))
def test_emulated_random_segments(emu, seed):
    """Seeded random cuts of in1 into 1..4 row segments, each identity or gathered (sorted or not, padded row stride or
    16-byte aligned), gradients stored / added / skipped: forward and both backward variants against the oracle."""
    rng = np.random.default_rng(90000 + seed)
    in2 = l2.sh_irreps(int(rng.integers(0, 3)))
    for _ in range(50):
        in1 = [(int(rng.integers(1, 12)), int(rng.integers(0, 3)), int(rng.choice([-1, 1]))) for _ in range(int(rng.integers(1, 7)))]
        out = [(int(rng.integers(1, 12)), int(rng.integers(0, 3)), int(rng.choice([-1, 1]))) for _ in range(int(rng.integers(1, 5)))]
        if l2.paths(in1, in2, out):
            break
    else:
        pytest.skip("no connected configuration drawn")
    rows, nn_ = int(rng.integers(1, 200)), int(rng.integers(1, 40))
    d1 = sum(m * (2 * l + 1) for m, l, _ in in1)
    d2 = sum(2 * l + 1 for _, l, _ in in2)
    do = sum(m * (2 * l + 1) for m, l, _ in out)
    nseg = int(rng.integers(1, min(4, d1) + 1))
    cuts = sorted(rng.choice(np.arange(1, d1), nseg - 1, replace=False).tolist()) if nseg > 1 else []
    widths = np.diff([0] + cuts + [d1]).tolist()
    parts, cols, modes, skip = [], [], [], []
    for wdt in widths:
        pad = int(rng.integers(0, 3))
        if rng.integers(0, 2):
            t = rng.standard_normal((nn_, wdt + pad)).astype(np.float32)
            idx = rng.integers(0, nn_, rows).astype(np.int32)
            srt = bool(rng.integers(0, 2))
            idx = np.sort(idx) if srt else idx
            parts.append((t, idx, wdt))
            cols.append(t[idx, :wdt])
            modes.append((3 if srt else 2) | (16 if wdt % 4 == 0 and (wdt + pad) % 4 == 0 else 0))
        else:
            t = rng.standard_normal((rows, wdt + pad)).astype(np.float32)
            parts.append((t, None, wdt))
            cols.append(t[:, :wdt])
            modes.append(1)
        skip.append(rng.integers(0, 5) == 0)
    modes = [0 if s else m for m, s in zip(modes, skip)] + [0] * (4 - nseg)
    x1 = np.ascontiguousarray(np.concatenate(cols, 1))
    y = rng.standard_normal((rows, d2)).astype(np.float32)
    nw = sum(a * b for a, b in l2.weight_shapes(in1, in2, out))
    w = rng.standard_normal(nw).astype(np.float32)
    g = rng.standard_normal((rows, do)).astype(np.float32)
    want_o, want_gx, want_gy, want_gw = _oracle(in1, in2, out, x1, y, w, g)
    spec = (len(in1), _flat(in1), len(in2), _flat(in2, True), len(out), _flat(out))
    got_o = np.full((rows, do), np.nan, np.float32)
    assert emu.emu_forward(*spec, C.c_longlong(rows), *Segs(parts).fwd(), _fp(y), _fp(w), _fp(got_o), 64, 256, 2) == 0
    _close(got_o, want_o)
    for fn in (emu.emu_backward, emu.emu_backward_split):
        gs = [None if s else (np.zeros_like(t) if ix is not None else np.full_like(t, np.nan)) for (t, ix, _), s in zip(parts, skip)]
        gy, gw = np.full((rows, d2), np.nan, np.float32), np.full(nw, np.nan, np.float32)
        rc = fn(*spec, C.c_longlong(rows), *Segs(parts, gs, modes).bwd(), _fp(y), _fp(w), _fp(g), _fp(gy), _fp(gw), 256, 2)
        if rc == 1 and fn is emu.emu_backward_split:
            continue
        assert rc == 0
        _close(gw, want_gw)
        _close(gy, want_gy)
        c0 = 0
        for (t, ix, wdt), gb in zip(parts, gs):
            ref = want_gx[:, c0:c0 + wdt]
            c0 += wdt
            if gb is None:
                continue
            if ix is None:
                _close(gb[:, :wdt], ref)
            else:
                acc = np.zeros((t.shape[0], wdt))
                np.add.at(acc, ix, ref)
                _close(gb[:, :wdt], acc)
                assert not gb[:, wdt:].any()
