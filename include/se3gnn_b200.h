/*
 * se3gnn_b200 — C ABI of the B200-native (sm_100a) hot path of Scalable-E3-GNN.
 *
 * Plain C, plain pointers and sizes; no torch types.  All device pointers are
 * CUDA device memory of the current device; `stream` is a cudaStream_t passed
 * as void*.  Every entry point returns 0 on success, a negative se3 error code
 * or a positive cudaError_t; se3_last_error() returns a message for the calling
 * thread.  There is NO CPU fallback anywhere behind this header.
 *
 * What each group replaces in the reference
 * (/root/reference/models/segnn/l1_tensor_prod.py, "L1TP"):
 *
 *   se3_l1tp_*      L1TensorProduct.__init__ species tables (L1TP:24-77) and
 *                   L1TensorProduct.forward (L1TP:234-299) + its autograd
 *                   backward (pure autograd in the reference).
 *   se3_gate_* (fused into se3_l1tp via the epilogue field), se3_edge_geom,
 *   se3_octree_*, se3_graph_*: the parts of the north-star path whose source is
 *                   not in the reference mount: builder-defined
 *                   spec, restated on CPU in oracle/.
 */
#ifndef SE3GNN_B200_H
#define SE3GNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SE3_OK 0
#define SE3_ERR_INVALID (-1)   /* bad argument / unsupported configuration   */
#define SE3_ERR_TOO_LARGE (-2) /* irreps do not fit the on-chip tiling        */
#define SE3_ERR_NO_DEVICE (-3) /* no sm_100 device / CUDA failure at init     */

#define SE3_MAX_SEG 4

const char* se3_last_error(void);
int se3_version(void);
/* number of kernels launched by this library in this process since load */
int64_t se3_launch_count(void);
/* of those, launches of the tcgen05 (tensor-core) kernels */
int64_t se3_tc_launch_count(void);

/* ---------------------------------------------------------------- l1tp ---- */

/* species order everywhere: 0 = 0e, 1 = 0o, 2 = 1e, 3 = 1o (L1TP:24-27) */
typedef struct se3_l1tp_desc {
    int32_t d_in1;             /* flat width of in1 (L1TP:20)                  */
    int32_t d_out;             /* flat width of out (L1TP:53)                  */
    int32_t n[4];              /* in1 channels per species (L1TP:67-73)        */
    int32_t m[4];              /* out channels per species (L1TP:74-77; l=1 /3)*/
    const int32_t* in_cols[4]; /* host: flat column of each channel (x comp.)  */
    const int32_t* out_cols[4];
} se3_l1tp_desc;

typedef struct se3_l1tp_plan se3_l1tp_plan; /* opaque */

int se3_l1tp_plan_create(const se3_l1tp_desc* desc, se3_l1tp_plan** plan);
void se3_l1tp_plan_destroy(se3_l1tp_plan* plan);
/* introspection for tests / bench: rows per tile, smem bytes, resident CTAs per SM */
int se3_l1tp_plan_info(const se3_l1tp_plan* plan, int32_t* tile_rows, int32_t* smem_fwd,
                       int32_t* smem_bwd, int32_t* weight_floats);

/* One segment of the (virtual) concatenation that forms an in1 row:
 * row r, columns [c0, c0+width) = base[(idx ? idx[r] : r) * ld + 0..width).   */
typedef struct se3_rowseg {
    const float* base;
    const int32_t* idx; /* NULL = identity */
    int32_t width;
    int32_t ld;
} se3_rowseg;

#define SE3_EPI_RAW 0  /* out = (f @ W) * norm                      (L1TP:250-297) */
#define SE3_EPI_GATE 1 /* + swish/sigmoid gate (O3TensorProductSwishGate, public SEGNN) */

typedef struct se3_l1tp_fwd_args {
    int64_t rows;
    int32_t nseg;
    se3_rowseg seg[SE3_MAX_SEG]; /* in1 = concat(seg...) ; sum(width) == d_in1 */
    const float* in2;            /* [rows,4] = (Y0, Y1x, Y1y, Y1z)  (L1TP:17)  */
    const float* w[4];           /* weights_l0e,l0o,l1e,l1o row-major (L1TP:81-88) or NULL */
    const float* norm[4];        /* norm_l0e,l0o,l1e,l1o (L1TP:159-162) or NULL (=1) */
    int32_t epilogue;            /* SE3_EPI_*                                   */
    int32_t gate_ns;             /* GATE: leading 0e outputs that are scalars; the
                                    remaining m0e-gate_ns (== m1o) 0e outputs gate the 1o outputs */
    float gate_cs, gate_cg;      /* normalize2mom constants for silu / sigmoid  */
    float* out_raw;              /* [rows,d_out] pre-activation (or the output in RAW mode); may be NULL */
    float* out_post;             /* GATE: [rows, gate_ns + 3*m1o]; may be NULL  */
    const float* resid;          /* RAW: added to the output, [rows,d_out]; may be NULL */
    const int32_t* seg_idx;      /* if non-NULL: the (post or raw) rows are summed into out_seg[seg_idx[r]].
                                    CONTRACT: seg_idx is GLOBALLY non-decreasing and out_seg is zero on entry
                                    (segments interior to a 64-row tile are written with a plain store; only
                                    tile-boundary segments accumulate) */
    float* out_seg;
} se3_l1tp_fwd_args;

int se3_l1tp_forward(se3_l1tp_plan* plan, const se3_l1tp_fwd_args* args, void* stream);

#define SE3_GRAD_NONE 0
#define SE3_GRAD_STORE 1  /* g[(idx?idx[r]:r)*ld + c]  = v   (identity rows)      */
#define SE3_GRAD_ATOMIC 2 /* g[idx[r]*ld + c]        += v   (unsorted gather)     */
#define SE3_GRAD_SORTED 3 /* same, idx sorted: run-length combined before the add */

typedef struct se3_l1tp_bwd_args {
    int64_t rows;
    int32_t nseg;
    se3_rowseg seg[SE3_MAX_SEG];
    const float* in2;
    const float* w[4];
    const float* norm[4];
    int32_t epilogue;
    int32_t gate_ns;
    float gate_cs, gate_cg;
    const float* raw;        /* GATE: saved pre-activation [rows,d_out]          */
    const float* gout;       /* cotangent of the forward result: [*, d_out] (RAW) or [*, d_post] (GATE) */
    const int32_t* gout_idx; /* row r reads gout[(gout_idx?gout_idx[r]:r)] — the transpose of seg_idx/out_seg */
    float* gseg[SE3_MAX_SEG];      /* gradient destination per in1 segment (same ld as forward) */
    int32_t gseg_mode[SE3_MAX_SEG];/* SE3_GRAD_*                                  */
    float* gw[4];            /* weight gradients (overwritten), NULL to skip all  */
    float* gin2;             /* [rows,4] or NULL (RAW epilogue only)              */
} se3_l1tp_bwd_args;

int se3_l1tp_backward(se3_l1tp_plan* plan, const se3_l1tp_bwd_args* args, void* stream);

/* ---------------------------------------------------------------- msg1 ---- */
/* First tensor product of the SEGNN message by linearity (csrc/msg_table.cu).  It replaces L1TensorProduct.forward
 * (L1TP:242-297) + its autograd backward for in1 = cat(x[dst], x[src], edge_extra), hidden irreps ns x0e + nv x1o,
 * output (ns+nv) x0e + nv x1o followed by the swish / sigmoid gate: the weight contraction runs once per NODE
 * (T = x . wbig, a dense [n_all, ns+3nv] x [ns+3nv, 8 (ns+2nv)] GEMM the caller performs), the per-edge work is the
 * combination with SH(1) and the gate.  Table / gradient rows: [dst half | src half], each [ns+2nv channels][4] =
 * (P, U_x, U_y, U_z), norms and 1/sqrt(3) folded in.  Weight layout: wz = weights_l0e [(2ns+2+2nv), ns+nv],
 * wv = weights_l1o [(2ns+2+2nv), nv], rows in the reference's concatenation order (L1TP:81-88); nz / nvn = norm_l0e /
 * norm_l1o or NULL. */
int se3_msg1_supported(int32_t ns, int32_t nv, int32_t n_extra); /* 1 if instantiated */
int se3_msg1_max_parts(void);                                   /* rows of gwe_part the backward may write */
int se3_msg1_expand(int32_t ns, int32_t nv, const float* wz, const float* wv, const float* nz, const float* nvn,
                    float* wbig /*[ns+3nv, 8(ns+2nv)]*/, float* we /*[2, ns+2nv]*/, void* stream);
/* one warp per destination node: rowptr [n_dst+1] = CSR of the edges by destination, src [E] (may point into the
 * halo rows of the table), y [E,4] = SH(1), extra [E,2]; writes pre [E, ns+4nv] (pre-activation, kept for backward)
 * and post [E, ns+3nv] (gated message). */
int se3_msg1_edge_forward(int32_t ns, int32_t nv, int64_t n_dst, const int64_t* rowptr, const int32_t* src,
                          const float* table /*[n_all, 8(ns+2nv)]*/, const float* we, const float* y,
                          const float* extra, float gate_cs, float gate_cg, float* pre, float* post, void* stream);
/* gate VJP + transposed SH combine + segment sums, no atomics: G [n_all, 8(ns+2nv)] is overwritten (dst half of the
 * rows >= n_dst: zero); tptr [n_all+1] / perm [E] = the edges in stable order by source (se3_graph_transpose);
 * gpre [E, ns+4nv] scratch (pre == NULL: gpre is an INPUT, the gate VJP was done by its producer and gpost is unused);
 * gwe_part [se3_msg1_max_parts(), 2, ns+2nv] per-block partials, *nparts rows written. */
int se3_msg1_edge_backward(int32_t ns, int32_t nv, int64_t n_dst, int64_t n_all, const int64_t* rowptr,
                           const int64_t* tptr, const int32_t* perm, const float* y, const float* extra,
                           const float* pre, const float* gpost, float gate_cs, float gate_cg, float* gpre, float* G,
                           float* gwe_part, int32_t* nparts, void* stream);
/* gwz / gwv (overwritten) from gwbig = x^T . G and the extras' partials */
int se3_msg1_contract(int32_t ns, int32_t nv, const float* gwbig, const float* gwe_part, int32_t nparts,
                      const float* nz, const float* nvn, float* gwz, float* gwv, void* stream);
/* The whole message layer forward in ONE launch (csrc/msg_fused_fwd.cu, tcgen05): node tables of message 1 -> SH
 * combine + gate -> weight contraction of message 2 on the tensor cores (3xTF32) -> SH combine + gate -> sorted-segment
 * sum over dst.  dst [rows] ascending; table / we as above (message 1); wz2 / wv2 / nz2 / nv2: weights_l0e
 * [(ns+nv), ns+nv], weights_l1o [(ns+nv), nv] and norms of message 2.  Writes pre1 [rows, ns+4nv], m1 [rows, ns+3nv],
 * pre2 [rows, ns+4nv] (what the backward reads) and agg [n_dst, ns+3nv] (+=: zero on entry).  pre1 / m1 / pre2 must be
 * ALLOCATED with the row count rounded up to a multiple of 64: whole 64-row tiles are written with cp.async.bulk (the
 * rows past `rows` receive unspecified values), and the backward kernels read whole tiles the same way. */
int se3_msg_fused_supported(int32_t ns, int32_t nv, int32_t n_extra);
int se3_msg_fused_forward(int32_t ns, int32_t nv, int64_t rows, const int32_t* dst, const int32_t* src,
                          const float* table, const float* we, const float* y, const float* extra, const float* wz2,
                          const float* wv2, const float* nz2, const float* nv2, float gate_cs, float gate_cg,
                          float* pre1, float* m1, float* pre2, float* agg, void* stream);
/* the same with cycle counters of the kernel's phases (diagnostics; dbg = NULL or [148][2][8] int64) */
int se3_msg_fused_forward_dbg(int32_t ns, int32_t nv, int64_t rows, const int32_t* dst, const int32_t* src,
                              const float* table, const float* we, const float* y, const float* extra, const float* wz2,
                              const float* wv2, const float* nz2, const float* nv2, float gate_cs, float gate_cg,
                              float* pre1, float* m1, float* pre2, float* agg, int64_t* dbg, void* stream);
/* Input-gradient side of the message layer in one launch (csrc/msg_fused_bwd.cu, tcgen05): cotangent of the aggregate
 * gagg [n_dst, ns+3nv] gathered through dst -> gate VJP of message 2 (pre2) -> contraction with W2^T (3xTF32) -> gate VJP
 * of message 1 (pre1) -> gpre1 [rows, ns+4nv], the input of se3_msg1_edge_backward(pre = NULL).  gpre2 (may be NULL)
 * receives the cotangent of message 2's pre-activation.  pre1 / pre2 / gpre1 / gpre2 must be ALLOCATED with the row
 * count rounded up to a multiple of 64 (whole tiles travel by cp.async.bulk in both directions). */
int se3_msg_fused_backward(int32_t ns, int32_t nv, int64_t rows, const int32_t* dst, const float* y, const float* pre1,
                           const float* pre2, const float* gagg, const float* wz2, const float* wv2, const float* nz2,
                           const float* nv2, float gate_cs, float gate_cg, float* gpre1, float* gpre2, void* stream);
/* Weight gradient of message 2 (csrc/msg_fused_bwdw.cu, tcgen05 MN-major, accumulators resident in TMEM): gwz2
 * [(ns+nv), ns+nv] and gwv2 [(ns+nv), nv] (overwritten) from the saved gated message 1 and gpre2 of
 * se3_msg_fused_backward; m1 / gpre2 must be allocated with the row count rounded up to 32 (whole tiles are copied by
 * cp.async.bulk; the extra rows may hold anything).  partials: max_parts x part_floats floats of scratch. */
int se3_msg_fused_bwdw_parts(int32_t ns, int32_t nv, int32_t* max_parts, int32_t* part_floats);
int se3_msg_fused_backward_w(int32_t ns, int32_t nv, int64_t rows, const float* y, const float* m1, const float* gpre2,
                             const float* nz2, const float* nv2, float* gwz2, float* gwv2, float* partials,
                             int32_t max_parts, void* stream);
/* The node-level contraction itself, block-sparse and in exact fp32 (csrc/msg_node.cu): table = x . W (forward),
 * gx = G . W^T and gwz / gwv = x^T . G (+ the extras' rows from gwe_part) straight from / into the parameters' layout;
 * se3_msg1_expand(wbig = NULL) then only produces `we`.  part: scratch of max_parts x part_floats floats
 * (se3_msg1_node_parts).  gx, or gwz and gwv together, may be NULL (skipped). */
int se3_msg1_node_parts(int32_t ns, int32_t nv, int32_t* max_parts, int32_t* part_floats);
int se3_msg1_node_table(int32_t ns, int32_t nv, int64_t n, const float* x, const float* wz, const float* wv,
                        const float* nz, const float* nvn, float* table, void* stream);
int se3_msg1_node_backward(int32_t ns, int32_t nv, int64_t n, const float* x, const float* G, const float* wz,
                           const float* wv, const float* nz, const float* nvn, const float* gwe_part, int32_t neparts,
                           float* gx, float* gwz, float* gwv, float* part, int32_t max_parts, void* stream);
/* -------------------------------------------------------------- domain ---- */
/* Morton-range domain decomposition of the replicated global graph (csrc/domain.cu; builder-defined, the reference
 * has no distributed code).  bounds [world+1] int64: particle-rank slab boundaries (device).  A cell belongs to the
 * owner of its first particle.
 * se3_domain_mark: pos [n+m] = local id of the nodes `rank` owns (-1 otherwise; owned particles first, then owned cells,
 *   both in global order), loc_rowptr [n_own+1] (capacity n+m+1) = CSR row pointers of the owned rows,
 *   counts [4+world] int64 (device): [0] owned particles, [1] owned nodes, [2] local edges.
 * se3_domain_edges (after reading counts[0..2]): own_ids [n_own], dst_loc / src_loc [e_loc] local ids in the global
 *   (dst, src) order (sources this rank does not own: n_own + position in halo_ids), the rank's rows of edge_attr
 *   [E,4] / edge_extra [E,2] (either may be NULL), halo_ids (capacity n+m; ascending global id inside each owner,
 *   grouped by owner), counts[3] = n_halo, counts[4+r] = halo nodes owned by rank r.
 * work: se3_domain_work_bytes(n+m); hpos [n+m], scratch [2 (n+m)] int32. */
int se3_domain_work_bytes(int64_t nn, size_t* bytes);
int se3_domain_mark(int64_t n, int64_t m, int32_t rank, int32_t world, const int64_t* bounds, const int32_t* cell_start,
                    const int64_t* rowptr, int32_t* pos, int64_t* loc_rowptr, int64_t* counts, void* work,
                    size_t work_bytes, void* stream);
int se3_domain_edges(int64_t n, int64_t m, int32_t world, const int64_t* bounds, const int32_t* cell_start,
                     const int32_t* pos, const int64_t* loc_rowptr, const int64_t* rowptr, const int32_t* col,
                     const float* edge_attr, const float* edge_extra, int64_t n_own, int64_t e_loc, int32_t* own_ids,
                     int32_t* dst_loc, int32_t* src_loc, float* attr_loc, float* extra_loc, int32_t* halo_ids,
                     int32_t* hpos, int32_t* scratch, int64_t* counts, void* work, size_t work_bytes, void* stream);
/* rowptr [n+1] of an ascending index (rowptr[k] = first position with idx >= k) */
int se3_rowptr_from_sorted(int64_t e, int64_t n, const int32_t* idx_sorted, int64_t* rowptr, void* stream);
/* stable counting sort of the edges by source: tptr [n_src+1], perm [e] (edge ids, ascending inside a segment) */
int se3_graph_transpose_work_bytes(int64_t n_src, size_t* bytes);
int se3_graph_transpose(int64_t e, int64_t n_src, const int32_t* src, int64_t* tptr, int32_t* perm, void* work,
                        size_t work_bytes, void* stream);

/* ---------------------------------------------------------------- o3tp ---- */
/* Fully connected O(3) tensor product for 0 <= l <= 2 (BASELINE configs[2]): the generalisation of
 * L1TensorProduct the reference excludes (L1TP:13-14 asserts lmax == 1), same conventions: paths enumerated
 * (i_out, i_in2, i_in1) (L1TP:122-151), 'component' x 'element' normalisation (L1TP:124,145,169), unit-norm couplings
 * that reduce to cg000/cg110/cg011/cg111 (L1TP:91-94) for l <= 1.  Specification: oracle/lmax2_oracle.py.
 * in2 is a spherical-harmonics type input (multiplicity 1 per irrep).  Flat layouts are e3nn's: a `mul x l` block is
 * [mul, 2l+1] row-major; l=1 components (x,y,z); l=2 components (xy, yz, 2zz-xx-yy, zx, xx-yy) (orthonormal).
 * Weights: one flat fp32 buffer, path p owns a [mul_in1, mul_out] row-major block at weight_offset[p]. */
#define SE3_O3_MAX_IRREPS 8
typedef struct se3_o3tp_desc {
    int32_t n_in1, n_in2, n_out;
    int32_t in1_mul[SE3_O3_MAX_IRREPS], in1_l[SE3_O3_MAX_IRREPS], in1_p[SE3_O3_MAX_IRREPS]; /* p = +1 / -1 */
    int32_t in2_l[3], in2_p[3];
    int32_t out_mul[SE3_O3_MAX_IRREPS], out_l[SE3_O3_MAX_IRREPS], out_p[SE3_O3_MAX_IRREPS];
} se3_o3tp_desc;

typedef struct se3_o3tp_plan se3_o3tp_plan; /* opaque */

int se3_o3tp_plan_create(const se3_o3tp_desc* desc, se3_o3tp_plan** plan);
void se3_o3tp_plan_destroy(se3_o3tp_plan* plan);
/* dims[0..7] = d_in1, d_in2, d_out, n_paths, weight_floats, rows per tile forward, rows per tile backward, backward shared memory KiB | accumulators-in-global flag << 16 | double-buffer flag << 17 | split-backward flag << 18 | weight gradient on the tensor cores (dense in1, whole 32-row tiles) << 19 | scalar second input handled as per-irrep linear maps << 20 */
int se3_o3tp_plan_info(const se3_o3tp_plan* plan, int32_t dims[8]);
/* host arrays of n_paths entries each (any may be NULL); path_weight = the normalisation factor a_out */
int se3_o3tp_plan_paths(const se3_o3tp_plan* plan, int32_t* i_in1, int32_t* i_in2, int32_t* i_out,
                        int32_t* weight_offset, float* path_weight);
/* host: the unit-norm coupling tensor of l1 x l2 -> l3 as [2l1+1][2l2+1][2l3+1] doubles; returns SE3_ERR_INVALID if
 * the triangle rule excludes the triple */
int se3_o3tp_coupling(int32_t l1, int32_t l2, int32_t l3, double* out);
/* out[rows, d_out] = TP(in1[rows, d_in1], in2[rows, d_in2]; w) */
int se3_o3tp_forward(se3_o3tp_plan* plan, int64_t rows, const float* in1, const float* in2, const float* w,
                     float* out, void* stream);
/* gin1[rows, d_in1] and gw[weight_floats] are overwritten; gin2[rows, d_in2] may be NULL (skipped) */
int se3_o3tp_backward(se3_o3tp_plan* plan, int64_t rows, const float* in1, const float* in2, const float* w,
                      const float* gout, float* gin1, float* gin2, float* gw, void* stream);

/* The same with in1 given as a virtual concatenation of up to SE3_MAX_SEG (optionally gathered) row segments, as in
 * se3_l1tp_fwd_args: the message product reads cat(x[dst], x[src], edge_extra) without materialising it.  Backward:
 * gseg[s] receives the gradient of segment s with gseg_mode[s] = SE3_GRAD_NONE (or gseg[s] == NULL) / SE3_GRAD_STORE
 * (identity rows: written) / SE3_GRAD_ATOMIC (gathered rows: += with atomics, caller zeroes) / SE3_GRAD_SORTED (the same
 * for an index that is sorted: runs inside a tile are summed before the one atomic add). */
int se3_o3tp_forward_seg(se3_o3tp_plan* plan, int64_t rows, int32_t nseg, const se3_rowseg* seg, const float* in2,
                         const float* w, float* out, void* stream);
int se3_o3tp_backward_seg(se3_o3tp_plan* plan, int64_t rows, int32_t nseg, const se3_rowseg* seg, const float* in2,
                          const float* w, const float* gout, float* const* gseg, const int32_t* gseg_mode, float* gin2,
                          float* gw, void* stream);

/* ---------------------------------------------------------------- o3msg ---- */
/* First tensor product of the l <= 2 SEGNN message layer by linearity (the op chain it replaces: the tensor product
 * above applied to cat(x[dst], x[src], extra), i.e. L1TP:242-297 generalised to l = 2; l <= 1 counterpart: se3_msg1_*).
 * The weight contraction runs once per node into per-role tables (made with se3_o3tp_forward on a scalar second input,
 * see se3gnn_b200/o3msg.py); per edge: pre[e][off + w (2l+1) + c] = a sum_paths sum_i M_p(Y_e)[i][c] (tdst[dst e] +
 * tsrc[src e])[tbase_p + w (2 l1 + 1) + i] + the scalar extras' paths (weights read from the flat buffer w).
 * One descriptor per output irrep; <= 4 table paths and <= 2 extras paths each. */
#define SE3_O3MSG_MAXP 4
#define SE3_O3MSG_MAXX 2
typedef struct se3_o3msg_io {
    int32_t l, mul, off;                    /* degree, multiplicity, first column of the output irrep in a pre row  */
    float a;                                /* normalisation factor of the output irrep (se3_o3tp_plan_paths)       */
    int32_t np;                             /* table paths into this irrep                                           */
    int32_t p_l1[SE3_O3MSG_MAXP], p_l2[SE3_O3MSG_MAXP], p_yoff[SE3_O3MSG_MAXP], p_tbase[SE3_O3MSG_MAXP];
    int32_t nx;                             /* extras paths (scalar extras x Y_l -> l)                               */
    int32_t x_l2[SE3_O3MSG_MAXX], x_yoff[SE3_O3MSG_MAXX], x_woff[SE3_O3MSG_MAXX] /* [x_mul, mul] block in w */,
        x_off[SE3_O3MSG_MAXX] /* first column in an extras row */, x_mul[SE3_O3MSG_MAXX];
    int32_t gx_off, gx_slots;               /* backward: gex[n][gx_off + w gx_slots + s], gx_slots = sum x_mul (<= 4) */
} se3_o3msg_io;
/* dst/src [edges] int32; tdst [*, ldt], tsrc [*, ldt]; y [edges, ldy]; extra [edges, ldx]; pre [edges, ldo] */
int se3_o3msg_edge_forward(const se3_o3msg_io* io, int32_t nio, int64_t edges, const int32_t* dst, const int32_t* src,
                           const float* tdst, const float* tsrc, int32_t ldt, const float* y, int32_t ldy,
                           const float* extra, int32_t ldx, const float* w, float* pre, int32_t ldo, void* stream);
/* Transpose: gdst [n_dst, ldt] = sums over the CSR rows (rowptr, edges sorted by destination), gsrc [n_all, ldt] = sums
 * over the transposed order (tptr, perm: se3_graph_transpose); every table entry is written (no zero-init, no atomics).
 * gex [n_dst, ldg]: per-destination partial sums of the extras' weight gradients (the caller adds the rows up). */
int se3_o3msg_edge_backward(const se3_o3msg_io* io, int32_t nio, int64_t n_dst, int64_t n_all, const int64_t* rowptr,
                            const int64_t* tptr, const int32_t* perm, const float* y, int32_t ldy, const float* extra,
                            int32_t ldx, const float* gpre, int32_t ldo, float* gdst, float* gsrc, int32_t ldt, float* gex,
                            int32_t ldg, void* stream);

/* ---------------------------------------------------------------- gate ---- */
/* Gated non-linearity on flat rows (public SEGNN O3TensorProductSwishGate's Gate; the reference mount has no source for
 * it): raw = [ns scalars | ng gate scalars | block b: cnt[b] channels x dim[b] components ...], ng = sum cnt;
 * out = [cs silu(s) | raw * cg sigmoid(gate of the channel)].  For l <= 1 hidden irreps the tensor-product kernels fuse
 * this (SE3_EPI_GATE); this stand-alone pair serves layouts with l = 2 blocks.  nblk <= 4, host arrays cnt/dim. */
int se3_gate_forward(int64_t rows, int32_t ns, int32_t nblk, const int32_t* cnt, const int32_t* dim, float cs, float cg,
                     const float* raw /*[rows, ns+ng+sum cnt*dim]*/, float* out /*[rows, ns+sum cnt*dim]*/, void* stream);
int se3_gate_backward(int64_t rows, int32_t ns, int32_t nblk, const int32_t* cnt, const int32_t* dim, float cs, float cg,
                      const float* raw, const float* gout, float* graw, void* stream);

/* Gate followed by the aggregation over the destination (public SEGNN: message 2's gate, then scatter-add over dst), for
 * edges sorted by destination seg [rows]: out[seg[e]][:] += gate(raw[e][:]) in tiles of 64 edges — no gated [E, d] tensor;
 * runs inside a tile are written with plain stores, only tile-boundary runs use atomic adds; out [n_seg, d_out] is
 * overwritten (zeroed first).  Backward: graw[e] = gate VJP(raw[e], gout[seg[e]]). */
int se3_gate_segment_sum_forward(int64_t rows, const int32_t* seg, int64_t n_seg, int32_t ns, int32_t nblk,
                                 const int32_t* cnt, const int32_t* dim, float cs, float cg, const float* raw, float* out,
                                 void* stream);
int se3_gate_segment_sum_backward(int64_t rows, const int32_t* seg, int32_t ns, int32_t nblk, const int32_t* cnt,
                                  const int32_t* dim, float cs, float cg, const float* raw, const float* gout /*[n_seg, d_out]*/,
                                  float* graw, void* stream);

/* -------------------------------------------------------------- octree ---- */
/* Builder-defined API (the reference's numba graph builder is not in the mount; the
 * specification is oracle/octree_oracle.py).  All arrays are caller-allocated device memory. */
typedef struct se3_octree {
    int64_t n;                 /* particles                                          */
    int32_t leaf_size;         /* a cell splits iff count > leaf_size                */
    int32_t max_depth;         /* 21 (63-bit Morton keys)                            */
    int64_t cell_cap;          /* capacity of the cell_* arrays                      */
    uint64_t* keys;            /* [n]  sorted Morton keys                            */
    int32_t* order;            /* [n]  rank -> original particle index               */
    int32_t* cell_start;       /* [cell_cap] first rank of the cell                  */
    int32_t* cell_count;       /* [cell_cap] particles in the cell                   */
    int32_t* cell_level;       /* [cell_cap]                                         */
    int32_t* cell_parent;      /* [cell_cap] -1 for the root                         */
    int32_t* cell_first_child; /* [cell_cap] -1 for leaves                           */
    int32_t* cell_nchild;      /* [cell_cap]                                         */
    uint64_t* cell_key;        /* [cell_cap] Morton prefix (key >> 3*(21-level))     */
    int32_t* level_ptr;        /* [max_depth+2] first cell id of each level          */
    int32_t* leaf_of_rank;     /* [n]  leaf cell of each rank                        */
    int32_t* cell_of_particle; /* [n]  leaf cell of each particle, original order    */
    float* bbox;               /* [4]  lo.x lo.y lo.z scale                          */
    void* work;                /* scratch, >= se3_octree_work_bytes()                */
    size_t work_bytes;
} se3_octree;

int se3_octree_work_bytes(int64_t n, int64_t cell_cap, size_t* bytes);
/* keys + stable radix sort + level-synchronous split + leaf assignment.  Synchronises the
 * stream once to return the number of cells and of non-empty levels. */
int se3_octree_build(const float* pos /*[n,3]*/, se3_octree* t, int64_t* m_out, int32_t* nlevels_out, void* stream);
/* 26-neighbour lookup + degrees + exclusive scan.  nbr [m,26], deg [n+m], rowptr [n+m+1],
 * scan_work [(n+m)/1024+2].  Synchronises once to return the edge count. */
int se3_graph_degrees(const se3_octree* t, int64_t m, int32_t* nbr, int32_t* deg, int64_t* rowptr,
                      int64_t* scan_work, int64_t* e_out, void* stream);
/* CSR emission: col [e] = source node, dst [e] = target node (nodes: ranks 0..n-1, cells n..n+m-1). */
int se3_graph_emit(const se3_octree* t, int64_t m, const int32_t* nbr, const int64_t* rowptr, int32_t* col,
                   int32_t* dst, void* stream);
/* node positions / velocities / masses in node order: particles permuted to rank order, cells = moments. */
int se3_node_data(const se3_octree* t, int64_t m, int32_t nlevels, const float* pos, const float* vel,
                  const float* mass, float* npos, float* nvel, float* nmass, void* stream);
/* edge_attr [e,4] = SH(1)(pos[src]-pos[dst]) ('integral' normalisation), edge_extra [e,2] = (|rel|, m_i m_j),
 * node_attr [n+m,4] = mean incoming edge_attr + SH(1)(vel), x_in [n+m,8] = (pos-centroid, vel, |vel|, mass). */
int se3_edge_geometry(int64_t n, int64_t m, int64_t e, const int64_t* rowptr, const int32_t* col,
                      const int32_t* dst, const float* npos, const float* nvel, const float* nmass,
                      float mass_scale, float* edge_attr, float* edge_extra, float* node_attr, float* x_in,
                      void* stream);

/* SH(2) attributes for the l <= 2 tensor product: edge_attr9 [e,9] = SH(2)(pos[src]-pos[dst]), columns Y0 | Y1 (x,y,z) |
 * Y2 (xy, yz, 2zz-xx-yy, zx, xx-yy), 'integral' normalisation (the first four columns are se3_edge_geometry's);
 * node_attr9 [n+m,9] = mean incoming edge_attr9 + SH(2)(vel). */
int se3_edge_geometry_l2(int64_t n, int64_t m, int64_t e, const int64_t* rowptr, const int32_t* col,
                         const int32_t* dst, const float* npos, const float* nvel, float* edge_attr9,
                         float* node_attr9, void* stream);

#ifdef __cplusplus
}
#endif
#endif
