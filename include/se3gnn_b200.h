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
    const int32_t* seg_idx;      /* if non-NULL: rows are sorted by seg_idx and the (post or raw)
                                    rows are summed into out_seg[seg_idx[r]] (+=, caller zeroes) */
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

#ifdef __cplusplus
}
#endif
#endif
