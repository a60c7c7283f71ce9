"""ctypes binding of ``include/se3gnn_b200.h`` (the C ABI of the CUDA library).

There is no CPU fallback: if the shared library is missing, or a call fails,
this module raises.  PyTorch is used by callers only for device memory and
streams; pointers cross this boundary as plain integers.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libse3gnn_b200.so")

MAX_SEG = 4
EPI_RAW, EPI_GATE = 0, 1
GRAD_NONE, GRAD_STORE, GRAD_ATOMIC, GRAD_SORTED = 0, 1, 2, 3

_i32p = C.POINTER(C.c_int32)
_f32p = C.c_void_p  # device pointers travel as integers


class L1tpDesc(C.Structure):
    _fields_ = [
        ("d_in1", C.c_int32), ("d_out", C.c_int32),
        ("n", C.c_int32 * 4), ("m", C.c_int32 * 4),
        ("in_cols", _i32p * 4), ("out_cols", _i32p * 4),
    ]


class RowSeg(C.Structure):
    _fields_ = [("base", C.c_void_p), ("idx", C.c_void_p), ("width", C.c_int32), ("ld", C.c_int32)]


O3MSG_MAXP, O3MSG_MAXX = 4, 2


class O3MsgIO(C.Structure):
    """``se3_o3msg_io``: one output irrep of the message product by linearity (csrc/o3msg.cu)."""
    _fields_ = [
        ("l", C.c_int32), ("mul", C.c_int32), ("off", C.c_int32), ("a", C.c_float), ("np", C.c_int32),
        ("p_l1", C.c_int32 * O3MSG_MAXP), ("p_l2", C.c_int32 * O3MSG_MAXP), ("p_yoff", C.c_int32 * O3MSG_MAXP),
        ("p_tbase", C.c_int32 * O3MSG_MAXP), ("nx", C.c_int32),
        ("x_l2", C.c_int32 * O3MSG_MAXX), ("x_yoff", C.c_int32 * O3MSG_MAXX), ("x_woff", C.c_int32 * O3MSG_MAXX),
        ("x_off", C.c_int32 * O3MSG_MAXX), ("x_mul", C.c_int32 * O3MSG_MAXX),
        ("gx_off", C.c_int32), ("gx_slots", C.c_int32),
    ]


class L1tpFwdArgs(C.Structure):
    _fields_ = [
        ("rows", C.c_int64), ("nseg", C.c_int32), ("seg", RowSeg * MAX_SEG),
        ("in2", C.c_void_p), ("w", C.c_void_p * 4), ("norm", C.c_void_p * 4),
        ("epilogue", C.c_int32), ("gate_ns", C.c_int32), ("gate_cs", C.c_float), ("gate_cg", C.c_float),
        ("out_raw", C.c_void_p), ("out_post", C.c_void_p), ("resid", C.c_void_p),
        ("seg_idx", C.c_void_p), ("out_seg", C.c_void_p),
    ]


class L1tpBwdArgs(C.Structure):
    _fields_ = [
        ("rows", C.c_int64), ("nseg", C.c_int32), ("seg", RowSeg * MAX_SEG),
        ("in2", C.c_void_p), ("w", C.c_void_p * 4), ("norm", C.c_void_p * 4),
        ("epilogue", C.c_int32), ("gate_ns", C.c_int32), ("gate_cs", C.c_float), ("gate_cg", C.c_float),
        ("raw", C.c_void_p), ("gout", C.c_void_p), ("gout_idx", C.c_void_p),
        ("gseg", C.c_void_p * MAX_SEG), ("gseg_mode", C.c_int32 * MAX_SEG),
        ("gw", C.c_void_p * 4), ("gin2", C.c_void_p),
    ]


class Octree(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("leaf_size", C.c_int32), ("max_depth", C.c_int32), ("cell_cap", C.c_int64),
        ("keys", C.c_void_p), ("order", C.c_void_p),
        ("cell_start", C.c_void_p), ("cell_count", C.c_void_p), ("cell_level", C.c_void_p),
        ("cell_parent", C.c_void_p), ("cell_first_child", C.c_void_p), ("cell_nchild", C.c_void_p),
        ("cell_key", C.c_void_p), ("level_ptr", C.c_void_p), ("leaf_of_rank", C.c_void_p),
        ("cell_of_particle", C.c_void_p), ("bbox", C.c_void_p), ("work", C.c_void_p), ("work_bytes", C.c_size_t),
    ]


O3_MAX_IRREPS = 8


class O3tpDesc(C.Structure):
    _fields_ = [
        ("n_in1", C.c_int32), ("n_in2", C.c_int32), ("n_out", C.c_int32),
        ("in1_mul", C.c_int32 * O3_MAX_IRREPS), ("in1_l", C.c_int32 * O3_MAX_IRREPS), ("in1_p", C.c_int32 * O3_MAX_IRREPS),
        ("in2_l", C.c_int32 * 3), ("in2_p", C.c_int32 * 3),
        ("out_mul", C.c_int32 * O3_MAX_IRREPS), ("out_l", C.c_int32 * O3_MAX_IRREPS), ("out_p", C.c_int32 * O3_MAX_IRREPS),
    ]


EXPORTS = [
    # name, restype, argtypes  (must list every symbol include/se3gnn_b200.h declares)
    ("se3_last_error", C.c_char_p, []),
    ("se3_version", C.c_int, []),
    ("se3_launch_count", C.c_int64, []),
    ("se3_tc_launch_count", C.c_int64, []),
    ("se3_l1tp_plan_create", C.c_int, [C.POINTER(L1tpDesc), C.POINTER(C.c_void_p)]),
    ("se3_l1tp_plan_destroy", None, [C.c_void_p]),
    ("se3_l1tp_plan_info", C.c_int, [C.c_void_p, _i32p, _i32p, _i32p, _i32p]),
    ("se3_l1tp_forward", C.c_int, [C.c_void_p, C.POINTER(L1tpFwdArgs), C.c_void_p]),
    ("se3_l1tp_backward", C.c_int, [C.c_void_p, C.POINTER(L1tpBwdArgs), C.c_void_p]),
    ("se3_msg1_supported", C.c_int, [C.c_int32, C.c_int32, C.c_int32]),
    ("se3_msg1_max_parts", C.c_int, []),
    ("se3_msg1_expand", C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    ("se3_msg1_edge_forward", C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("se3_msg1_edge_backward", C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                         C.c_void_p, C.c_void_p, C.c_void_p, _i32p, C.c_void_p]),
    ("se3_msg1_contract", C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    ("se3_msg_fused_supported", C.c_int, [C.c_int32, C.c_int32, C.c_int32]),
    ("se3_msg_fused_forward", C.c_int, [C.c_int32, C.c_int32, C.c_int64] + [C.c_void_p] * 10 + [C.c_float, C.c_float]
     + [C.c_void_p] * 5),
    ("se3_msg_fused_bwdw_parts", C.c_int, [C.c_int32, C.c_int32, _i32p, _i32p]),
    ("se3_msg_fused_backward_w", C.c_int, [C.c_int32, C.c_int32, C.c_int64] + [C.c_void_p] * 8 + [C.c_int32, C.c_void_p]),
    ("se3_msg1_node_parts", C.c_int, [C.c_int32, C.c_int32, _i32p, _i32p]),
    ("se3_msg1_node_table", C.c_int, [C.c_int32, C.c_int32, C.c_int64] + [C.c_void_p] * 7),
    ("se3_msg1_node_backward", C.c_int, [C.c_int32, C.c_int32, C.c_int64] + [C.c_void_p] * 7 + [C.c_int32]
     + [C.c_void_p] * 4 + [C.c_int32, C.c_void_p]),
    ("se3_msg_fused_forward_dbg", C.c_int, [C.c_int32, C.c_int32, C.c_int64] + [C.c_void_p] * 10 + [C.c_float, C.c_float]
     + [C.c_void_p] * 6),
    ("se3_msg_fused_backward", C.c_int, [C.c_int32, C.c_int32, C.c_int64] + [C.c_void_p] * 9 + [C.c_float, C.c_float]
     + [C.c_void_p] * 3),
    ("se3_domain_work_bytes", C.c_int, [C.c_int64, C.POINTER(C.c_size_t)]),
    ("se3_domain_mark", C.c_int, [C.c_int64, C.c_int64, C.c_int32, C.c_int32] + [C.c_void_p] * 7 + [C.c_size_t, C.c_void_p]),
    ("se3_domain_edges", C.c_int, [C.c_int64, C.c_int64, C.c_int32] + [C.c_void_p] * 8 + [C.c_int64, C.c_int64]
     + [C.c_void_p] * 10 + [C.c_size_t, C.c_void_p]),
    ("se3_rowptr_from_sorted", C.c_int, [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("se3_graph_transpose_work_bytes", C.c_int, [C.c_int64, C.POINTER(C.c_size_t)]),
    ("se3_graph_transpose", C.c_int, [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                      C.c_void_p]),
    ("se3_o3tp_plan_create", C.c_int, [C.POINTER(O3tpDesc), C.POINTER(C.c_void_p)]),
    ("se3_o3tp_plan_destroy", None, [C.c_void_p]),
    ("se3_o3tp_plan_info", C.c_int, [C.c_void_p, _i32p]),
    ("se3_o3tp_plan_paths", C.c_int, [C.c_void_p, _i32p, _i32p, _i32p, _i32p, C.POINTER(C.c_float)]),
    ("se3_o3tp_coupling", C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double)]),
    ("se3_o3tp_forward", C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("se3_o3tp_backward", C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    ("se3_o3tp_forward_seg", C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(RowSeg), C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    ("se3_o3tp_backward_seg", C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(RowSeg), C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.POINTER(C.c_void_p), _i32p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("se3_o3msg_edge_forward", C.c_int, [C.POINTER(O3MsgIO), C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                         C.c_void_p, C.c_int32, C.c_void_p]),
    ("se3_o3msg_edge_backward", C.c_int, [C.POINTER(O3MsgIO), C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                          C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    ("se3_gate_forward", C.c_int, [C.c_int64, C.c_int32, C.c_int32, _i32p, _i32p, C.c_float, C.c_float, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    ("se3_gate_backward", C.c_int, [C.c_int64, C.c_int32, C.c_int32, _i32p, _i32p, C.c_float, C.c_float, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    ("se3_gate_segment_sum_forward", C.c_int, [C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, _i32p, _i32p, C.c_float,
                                               C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("se3_gate_segment_sum_backward", C.c_int, [C.c_int64, C.c_void_p, C.c_int32, C.c_int32, _i32p, _i32p, C.c_float,
                                                C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("se3_octree_work_bytes", C.c_int, [C.c_int64, C.c_int64, C.POINTER(C.c_size_t)]),
    ("se3_octree_build", C.c_int, [C.c_void_p, C.POINTER(Octree), C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.c_void_p]),
    ("se3_graph_degrees", C.c_int, [C.POINTER(Octree), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.POINTER(C.c_int64), C.c_void_p]),
    ("se3_graph_emit", C.c_int, [C.POINTER(Octree), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("se3_node_data", C.c_int, [C.POINTER(Octree), C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("se3_edge_geometry", C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    ("se3_edge_geometry_l2", C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
]

_lib = None


class Se3Error(RuntimeError):
    pass


def lib():
    """Load the CUDA library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Se3Error(
                f"{LIB_PATH} is missing: build it with `python -m se3gnn_b200.build` "
                "(nvcc, sm_100a).  se3gnn_b200 has no CPU / eager fallback.")
        L = C.CDLL(LIB_PATH)
        for name, res, args in EXPORTS:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str = "se3gnn_b200"):
    if rc != 0:
        msg = lib().se3_last_error().decode(errors="replace")
        raise Se3Error(f"{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(lib().se3_launch_count())


def tc_launch_count() -> int:
    return int(lib().se3_tc_launch_count())


# ---------------------------------------------------------------- per-launch profiling (bench.py)
_prof = None


def profile_begin():
    """Start recording one (tag, CUDA-event pair, algorithmic bytes, flops) entry per library call."""
    global _prof
    _prof = []


def profile_end():
    """Stop recording; returns [(tag, ms, bytes, flops)] (synchronises)."""
    global _prof
    import torch
    torch.cuda.synchronize()
    out = [(tag, a.elapsed_time(b), nbytes, flops) for tag, a, b, nbytes, flops in (_prof or [])]
    _prof = None
    return out


class mark:
    """Context manager: CUDA events on the launching (current) stream around one library call."""

    def __init__(self, tag: str, nbytes: float = 0.0, flops: float = 0.0):
        self.tag, self.nbytes, self.flops = tag, nbytes, flops

    def __enter__(self):
        if _prof is not None:
            import torch
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if _prof is not None:
            self.b.record()
            _prof.append((self.tag, self.a, self.b, self.nbytes, self.flops))
        return False


def ptr(t) -> Optional[int]:
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    return t.data_ptr()


def current_stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


class L1tpPlan:
    """Owns a ``se3_l1tp_plan`` (device-side column tables + tiling)."""

    def __init__(self, n: Sequence[int], m: Sequence[int], in_cols, out_cols):
        self.n = [int(x) for x in n]
        self.m = [int(x) for x in m]
        d = L1tpDesc()
        d.d_in1 = self.n[0] + self.n[1] + 3 * (self.n[2] + self.n[3])
        d.d_out = self.m[0] + self.m[1] + 3 * (self.m[2] + self.m[3])
        self.d_in1, self.d_out = d.d_in1, d.d_out
        self._keep = []
        for s in range(4):
            d.n[s] = self.n[s]
            d.m[s] = self.m[s]
            a = (C.c_int32 * max(1, self.n[s]))(*[int(c) for c in in_cols[s]])
            b = (C.c_int32 * max(1, self.m[s]))(*[int(c) for c in out_cols[s]])
            self._keep += [a, b]
            d.in_cols[s] = C.cast(a, _i32p)
            d.out_cols[s] = C.cast(b, _i32p)
        h = C.c_void_p()
        check(lib().se3_l1tp_plan_create(C.byref(d), C.byref(h)), "se3_l1tp_plan_create")
        self.handle = h

    def info(self):
        v = [C.c_int32() for _ in range(4)]
        check(lib().se3_l1tp_plan_info(self.handle, *[C.byref(x) for x in v]))
        return dict(tile_rows=v[0].value, smem_fwd=v[1].value, smem_bwd=v[2].value, weight_floats=v[3].value)

    def __del__(self):
        try:
            if getattr(self, "handle", None) and _lib is not None:
                _lib.se3_l1tp_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class O3tpPlan:
    """Owns a ``se3_o3tp_plan`` (l <= 2 tensor product: path list, coupling tables, tiling)."""

    def __init__(self, in1, in2, out):
        """in1/out: [(mul, l, p)], in2: [(l, p)] (multiplicity 1 each)."""
        d = O3tpDesc()
        if len(in1) > O3_MAX_IRREPS or len(out) > O3_MAX_IRREPS or len(in2) > 3:
            raise Se3Error(f"o3tp supports up to {O3_MAX_IRREPS} in1/out irreps and 3 in2 irreps")
        d.n_in1, d.n_in2, d.n_out = len(in1), len(in2), len(out)
        for i, (mul, l, p) in enumerate(in1):
            d.in1_mul[i], d.in1_l[i], d.in1_p[i] = int(mul), int(l), int(p)
        for i, (l, p) in enumerate(in2):
            d.in2_l[i], d.in2_p[i] = int(l), int(p)
        for i, (mul, l, p) in enumerate(out):
            d.out_mul[i], d.out_l[i], d.out_p[i] = int(mul), int(l), int(p)
        h = C.c_void_p()
        check(lib().se3_o3tp_plan_create(C.byref(d), C.byref(h)), "se3_o3tp_plan_create")
        self.handle = h
        dims = (C.c_int32 * 8)()
        check(lib().se3_o3tp_plan_info(h, dims))
        self.d_in1, self.d_in2, self.d_out, self.n_paths, self.weight_floats, self.tile_fwd, self.tile_bwd = list(dims)[:7]
        self.split_backward = bool((dims[7] >> 18) & 1)   # input / weight gradients as two kernels
        self.linear_maps = bool((dims[7] >> 20) & 1)      # scalar in2: forward / input gradients as linear maps (o3tp_lin.cu)
        self.tc_weight_grad = bool((dims[7] >> 19) & 1)   # weight gradient of a dense in1 on the tensor cores (tcgen05)
        n = max(1, self.n_paths)
        arrs = [(C.c_int32 * n)() for _ in range(4)]
        pw = (C.c_float * n)()
        check(lib().se3_o3tp_plan_paths(h, *arrs, pw))
        self.paths = [(arrs[0][k], arrs[1][k], arrs[2][k], arrs[3][k], float(pw[k])) for k in range(self.n_paths)]

    def __del__(self):
        try:
            if getattr(self, "handle", None) and _lib is not None:
                _lib.se3_o3tp_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def o3tp_coupling(l1: int, l2: int, l3: int):
    """Nested list [2l1+1][2l2+1][2l3+1] of the library's unit-norm coupling tensor."""
    n = (2 * l1 + 1) * (2 * l2 + 1) * (2 * l3 + 1)
    buf = (C.c_double * n)()
    check(lib().se3_o3tp_coupling(l1, l2, l3, buf), "se3_o3tp_coupling")
    it = iter(buf)
    return [[[next(it) for _ in range(2 * l3 + 1)] for _ in range(2 * l2 + 1)] for _ in range(2 * l1 + 1)]
