// First tensor product of the l <= 2 SEGNN message layer by linearity (BASELINE configs[2]; the l <= 1 counterpart is
// msg_table.cu).  TP(cat(x[dst], x[src], extra), Y) is linear in its first input for fixed spherical harmonics, and every
// path factorises into "weight contraction of the raw channels" followed by "coupling with Y":
//     out[io][w][c] = a_io sum_{paths p = (h, i2) into io} sum_i M_p(Y)[i][c] T[p][w][i],   M_p[i][c] = sum_j C[i][j][c] Y_{i2}[j]
//     T[p][w][i]    = sum_u W_p[u][w] x[u][i]                       (no Y in it: computed once per NODE, not per edge)
// so the contraction runs on ~16x fewer rows than there are edges (node tables, one per role dst / src, produced by the
// l <= 2 tensor-product operator itself with a scalar second input, see se3gnn_b200/o3msg.py) and the per-edge work is a
// gather of two table rows plus the coupling.  The backward is the exact transpose: coupling^T of the cotangent per
// edge, summed over the CSR row of a destination / the transposed-order row of a source — one thread per (node, output
// channel), no atomics, run-to-run deterministic.
//
// Table row of a role (floats): for every hidden irrep h (degree l1) a block of N_h columns, column col = first column
// of the path + w, entry (col, i) at base_h + col (2 l1 + 1) + i; `tbase` of a path = base_h + first column * (2 l1 + 1).
// Specification: oracle/lmax2_oracle.py (same couplings and normalisation as csrc/o3tp.cu: o3tp_cg_gen.inl).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace {

using se3::set_error;

#define O3_DEV __device__ __forceinline__
#include "o3tp_cg_gen.inl"

constexpr int NTH = 256;

struct EdgeArgs {
    se3_o3msg_io io;
    long long E;
    const int32_t* dst;
    const int32_t* src;
    const float* tdst;
    const float* tsrc;
    int ldt;
    const float* y;
    int ldy;
    const float* ex;
    int ldx;
    const float* w;
    float* pre;
    int ldo;
    int per_block;      // edges (nodes) per block = NTH / mul
    unsigned magic;     // ceil(2^16 / mul): tid / mul = (tid * magic) >> 16 for tid < 256
    int ny;             // in2 is SH-type with Y_l at column l^2 (checked by the host): the ny <= 9 values of a row are read
                        // once into registers; 0: general offsets, every path reads its Y through the pointer
};

struct NodeArgs {
    se3_o3msg_io io;
    long long n;
    const long long* ptr;
    const int32_t* perm;
    const float* y;
    int ldy;
    const float* ex;
    int ldx;
    const float* gpre;
    int ldo;
    float* g;
    int ldt;
    float* gex;
    int ldg;
    int per_block;
    unsigned magic;
    int ny;
};

template <int L1, int L2, int LO>
O3_DEV void fwd_path(const float* __restrict__ td, const float* __restrict__ ts, const float* __restrict__ yr,
                     float (&acc)[2 * LO + 1]) {
    constexpr int D1 = 2 * L1 + 1, DO = 2 * LO + 1;
    float M[D1][DO];
    o3_M<L1, L2, LO>(yr, M);
#pragma unroll
    for (int i = 0; i < D1; ++i) {
        const float t = __ldg(td + i) + __ldg(ts + i);
#pragma unroll
        for (int c = 0; c < DO; ++c)
            if ((o3_nz<L1, L2, LO>::mask >> (i * DO + c)) & 1u) acc[c] = fmaf(M[i][c], t, acc[c]);
    }
}

template <int L2, int LO>
O3_DEV void fwd_scalar(float t, const float* __restrict__ yr, float (&acc)[2 * LO + 1]) {
    constexpr int DO = 2 * LO + 1;
    float M[1][DO];
    o3_M<0, L2, LO>(yr, M);
#pragma unroll
    for (int c = 0; c < DO; ++c) acc[c] = fmaf(M[0][c], t, acc[c]);
}

template <int L1, int L2, int LO>
O3_DEV void bwd_path(const float* __restrict__ yr, const float (&g)[2 * LO + 1], float (&gt)[5]) {
    constexpr int D1 = 2 * L1 + 1, DO = 2 * LO + 1;
    float M[D1][DO];
    o3_M<L1, L2, LO>(yr, M);
#pragma unroll
    for (int i = 0; i < D1; ++i)
#pragma unroll
        for (int c = 0; c < DO; ++c)
            if ((o3_nz<L1, L2, LO>::mask >> (i * DO + c)) & 1u) gt[i] = fmaf(M[i][c], g[c], gt[i]);
}

// pre[e][off + w (2 LO + 1) + c] for one output irrep: thread = (edge, output channel w)
template <int LO, bool SH>
__global__ void __launch_bounds__(NTH) o3msg_edge_fwd_kernel(const __grid_constant__ EdgeArgs A) {
    constexpr int DO = 2 * LO + 1;
    const int mul = A.io.mul;
    const int el = (int)((threadIdx.x * A.magic) >> 16), w = threadIdx.x - el * mul;
    if (el >= A.per_block) return;
    float wx[SE3_O3MSG_MAXX][4];   // this thread's extras weights (its output channel is fixed): out of the edge loop
#pragma unroll
    for (int p = 0; p < SE3_O3MSG_MAXX; ++p)
#pragma unroll
        for (int u = 0; u < 4; ++u)
            wx[p][u] = (p < A.io.nx && u < A.io.x_mul[p]) ? __ldg(A.w + A.io.x_woff[p] + u * mul + w) : 0.f;
    for (long long e = (long long)blockIdx.x * A.per_block + el; e < A.E; e += (long long)gridDim.x * A.per_block) {
        const float* td = A.tdst + (long long)__ldg(A.dst + e) * A.ldt;
        const float* ts = A.tsrc + (long long)__ldg(A.src + e) * A.ldt;
        const float* yr = A.y + e * A.ldy;
        float yreg[9];
        if (SH) {
#pragma unroll
            for (int j = 0; j < 9; ++j) yreg[j] = j < A.ny ? __ldg(yr + j) : 0.f;
        }
        float acc[DO];
#pragma unroll
        for (int c = 0; c < DO; ++c) acc[c] = 0.f;
#pragma unroll
        for (int p = 0; p < SE3_O3MSG_MAXP; ++p) {
            if (p < A.io.np) {
                const int l1 = A.io.p_l1[p], o = A.io.p_tbase[p] + w * (2 * l1 + 1);
                const float* yp = yr + A.io.p_yoff[p];
                switch (l1 * 9 + A.io.p_l2[p] * 3 + LO) {
#define O3M_CASE(a, b, c)                                                                         \
    case a * 9 + b * 3 + c:                                                                       \
        if constexpr (c == LO) fwd_path<a, b, c>(td + o, ts + o, SH ? yreg + b * b : yp, acc);    \
        break;
                    O3_TRIPLES(O3M_CASE)
#undef O3M_CASE
                }
            }
        }
#pragma unroll
        for (int p = 0; p < SE3_O3MSG_MAXX; ++p) {
            if (p < A.io.nx) {
                float t = 0.f;
                const float* xp = A.ex + e * A.ldx + A.io.x_off[p];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (u < A.io.x_mul[p]) t = fmaf(wx[p][u], __ldg(xp + u), t);
                const float* yp = yr + A.io.x_yoff[p];
                switch (A.io.x_l2[p]) {   // a scalar extra couples with Y_l into an output of the same degree
                    case 0: if constexpr (LO == 0) fwd_scalar<0, 0>(t, SH ? yreg : yp, acc); break;
                    case 1: if constexpr (LO == 1) fwd_scalar<1, 1>(t, SH ? yreg + 1 : yp, acc); break;
                    case 2: if constexpr (LO == 2) fwd_scalar<2, 2>(t, SH ? yreg + 4 : yp, acc); break;
                }
            }
        }
        float* o = A.pre + e * A.ldo + A.io.off + w * DO;
#pragma unroll
        for (int c = 0; c < DO; ++c) o[c] = A.io.a * acc[c];
    }
}

// G[n][tbase_p + w (2 l1 + 1) + i] = sum over the edges of node n of coupling^T (a g): thread = (node, output channel w).
// SRC = false: the CSR row of a destination (edges ptr[n] .. ptr[n+1]); also the extras' weight-gradient partials
// gex[n][gx_off + w slots + s] = sum_e extra[e][s] q_e.  SRC = true: edges perm[ptr[n] .. ptr[n+1]) of a source.
template <int LO, bool SRC, bool SH>
__global__ void __launch_bounds__(NTH, 4) o3msg_edge_bwd_kernel(const __grid_constant__ NodeArgs A) {
    constexpr int DO = 2 * LO + 1;
    const int mul = A.io.mul;
    const int nl = (int)((threadIdx.x * A.magic) >> 16), w = threadIdx.x - nl * mul;
    const long long n = (long long)blockIdx.x * A.per_block + nl;
    if (nl >= A.per_block || n >= A.n) return;
    float gt[SE3_O3MSG_MAXP][5];
    float sx[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int p = 0; p < SE3_O3MSG_MAXP; ++p)
#pragma unroll
        for (int i = 0; i < 5; ++i) gt[p][i] = 0.f;
    const long long beg = __ldg(A.ptr + n), end = __ldg(A.ptr + n + 1);
    // SH: the cotangent channel and the Y row of edge k + 1 are fetched while edge k is processed (the loop is serial per
    // thread and was latency-bound: long scoreboard 70-77 % of the samples)
    float gn[DO], yn[9];
    long long en = 0;
    auto fetch = [&](long long kk) {
        en = SRC ? (long long)__ldg(A.perm + kk) : kk;
        const float* gp = A.gpre + en * A.ldo + A.io.off + w * DO;
        const float* yq = A.y + en * A.ldy;
#pragma unroll
        for (int c = 0; c < DO; ++c) gn[c] = __ldg(gp + c);
#pragma unroll
        for (int j = 0; j < 9; ++j) yn[j] = j < A.ny ? __ldg(yq + j) : 0.f;
    };
    if (SH && beg < end) fetch(beg);
    for (long long k = beg; k < end; ++k) {
        long long e;
        float yreg[9], g[DO];
        const float* yr;
        if (SH) {
            e = en;
#pragma unroll
            for (int j = 0; j < 9; ++j) yreg[j] = yn[j];
#pragma unroll
            for (int c = 0; c < DO; ++c) g[c] = A.io.a * gn[c];
            yr = A.y + e * A.ldy;
            if (k + 1 < end) fetch(k + 1);
        } else {
            e = SRC ? (long long)__ldg(A.perm + k) : k;
            const float* gp = A.gpre + e * A.ldo + A.io.off + w * DO;
            yr = A.y + e * A.ldy;
#pragma unroll
            for (int c = 0; c < DO; ++c) g[c] = A.io.a * __ldg(gp + c);
        }
#pragma unroll
        for (int p = 0; p < SE3_O3MSG_MAXP; ++p) {
            if (p < A.io.np) {
                const float* yp = yr + A.io.p_yoff[p];
                switch (A.io.p_l1[p] * 9 + A.io.p_l2[p] * 3 + LO) {
#define O3M_CASE(a, b, c)                                                              \
    case a * 9 + b * 3 + c:                                                            \
        if constexpr (c == LO) bwd_path<a, b, c>(SH ? yreg + b * b : yp, g, gt[p]);    \
        break;
                    O3_TRIPLES(O3M_CASE)
#undef O3M_CASE
                }
            }
        }
        if (!SRC) {
            int s0 = 0;
#pragma unroll
            for (int p = 0; p < SE3_O3MSG_MAXX; ++p) {
                if (p < A.io.nx) {
                    float q[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
                    const float* yp = yr + A.io.x_yoff[p];
                    switch (A.io.x_l2[p]) {
                        case 0: if constexpr (LO == 0) bwd_path<0, 0, 0>(SH ? yreg : yp, g, q); break;
                        case 1: if constexpr (LO == 1) bwd_path<0, 1, 1>(SH ? yreg + 1 : yp, g, q); break;
                        case 2: if constexpr (LO == 2) bwd_path<0, 2, 2>(SH ? yreg + 4 : yp, g, q); break;
                    }
                    const float* xp = A.ex + e * A.ldx + A.io.x_off[p];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (u < A.io.x_mul[p] && s0 + u < 4) sx[s0 + u] = fmaf(__ldg(xp + u), q[0], sx[s0 + u]);
                    s0 += A.io.x_mul[p];
                }
            }
        }
    }
    float* gr = A.g + n * A.ldt;
#pragma unroll
    for (int p = 0; p < SE3_O3MSG_MAXP; ++p) {
        if (p < A.io.np) {
            const int d1 = 2 * A.io.p_l1[p] + 1;
            float* o = gr + A.io.p_tbase[p] + w * d1;
#pragma unroll
            for (int i = 0; i < 5; ++i)
                if (i < d1) o[i] = gt[p][i];
        }
    }
    if (!SRC && A.gex && A.io.gx_slots > 0) {
        float* o = A.gex + n * A.ldg + A.io.gx_off + w * A.io.gx_slots;
#pragma unroll
        for (int s = 0; s < 4; ++s)
            if (s < A.io.gx_slots) o[s] = sx[s];
    }
}

// in2 of SH type: Y_l at column l^2 for every path, and the rows hold all of them
int sh_width(const se3_o3msg_io* io, int nio, int ldy) {
    int lmax = 0;
    for (int k = 0; k < nio; ++k) {
        for (int p = 0; p < io[k].np; ++p) {
            if (io[k].p_yoff[p] != io[k].p_l2[p] * io[k].p_l2[p]) return 0;
            lmax = std::max(lmax, io[k].p_l2[p]);
        }
        for (int p = 0; p < io[k].nx; ++p) {
            if (io[k].x_yoff[p] != io[k].x_l2[p] * io[k].x_l2[p]) return 0;
            lmax = std::max(lmax, io[k].x_l2[p]);
        }
    }
    const int ny = (lmax + 1) * (lmax + 1);
    return ny <= ldy ? ny : 0;
}

int check_io(const se3_o3msg_io* io, int nio) {
    if (!io || nio < 1 || nio > 8) return 1;
    for (int k = 0; k < nio; ++k) {
        const se3_o3msg_io& I = io[k];
        if (I.l < 0 || I.l > 2 || I.mul < 1 || I.mul > NTH || I.np < 0 || I.np > SE3_O3MSG_MAXP || I.nx < 0 ||
            I.nx > SE3_O3MSG_MAXX || I.gx_slots < 0 || I.gx_slots > 4)
            return 1;
        int slots = 0;
        for (int p = 0; p < I.np; ++p)
            if (I.p_l1[p] < 0 || I.p_l1[p] > 2 || I.p_l2[p] < 0 || I.p_l2[p] > 2 || I.p_tbase[p] < 0 || I.p_yoff[p] < 0 ||
                I.l < std::abs(I.p_l1[p] - I.p_l2[p]) || I.l > I.p_l1[p] + I.p_l2[p])
                return 1;
        for (int p = 0; p < I.nx; ++p) {
            if (I.x_l2[p] != I.l || I.x_mul[p] < 1 || I.x_mul[p] > 4) return 1;
            slots += I.x_mul[p];
        }
        if (slots != I.gx_slots) return 1;
    }
    return 0;
}

}  // namespace

extern "C" int se3_o3msg_edge_forward(const se3_o3msg_io* io, int32_t nio, int64_t edges, const int32_t* dst,
                                      const int32_t* src, const float* tdst, const float* tsrc, int32_t ldt, const float* y,
                                      int32_t ldy, const float* extra, int32_t ldx, const float* w, float* pre, int32_t ldo,
                                      void* stream) {
    if (check_io(io, nio) || edges < 0 || ldt < 1 || ldy < 1 || ldo < 1) {
        set_error("o3msg forward: bad descriptor (<= 8 output irreps, <= %d table paths and <= %d extras paths each)",
                  SE3_O3MSG_MAXP, SE3_O3MSG_MAXX);
        return SE3_ERR_INVALID;
    }
    if (edges == 0) return SE3_OK;
    if (!dst || !src || !tdst || !tsrc || !y || !pre) { set_error("o3msg forward: null argument"); return SE3_ERR_INVALID; }
    for (int k = 0; k < nio; ++k) {
        if (io[k].nx > 0 && (!extra || !w)) { set_error("o3msg forward: extras paths need extra and w"); return SE3_ERR_INVALID; }
        EdgeArgs A;
        A.io = io[k]; A.E = edges; A.dst = dst; A.src = src; A.tdst = tdst; A.tsrc = tsrc; A.ldt = ldt; A.y = y; A.ldy = ldy;
        A.ex = extra; A.ldx = ldx; A.w = w; A.pre = pre; A.ldo = ldo;
        A.per_block = NTH / io[k].mul;
        A.magic = (65536u + io[k].mul - 1) / io[k].mul;
        const long long blocks = (edges + A.per_block - 1) / A.per_block;
        const int grid = (int)std::min<long long>(blocks, (long long)se3::num_sms() * 64);
        cudaStream_t st = (cudaStream_t)stream;
        // measured: reading Y through the pointer per path is faster in the forward (49.7 vs 58.0 ms per step at 17.9M
        // edges), the register copy wins in the backward (77 -> 64 ms), whose edge loop is serial per thread
        A.ny = getenv("SE3_O3MSG_FWD_YREG") ? sh_width(io, nio, ldy) : 0;
        if (A.ny) {
            switch (io[k].l) {
                case 0: o3msg_edge_fwd_kernel<0, true><<<grid, NTH, 0, st>>>(A); break;
                case 1: o3msg_edge_fwd_kernel<1, true><<<grid, NTH, 0, st>>>(A); break;
                default: o3msg_edge_fwd_kernel<2, true><<<grid, NTH, 0, st>>>(A); break;
            }
        } else {
            switch (io[k].l) {
                case 0: o3msg_edge_fwd_kernel<0, false><<<grid, NTH, 0, st>>>(A); break;
                case 1: o3msg_edge_fwd_kernel<1, false><<<grid, NTH, 0, st>>>(A); break;
                default: o3msg_edge_fwd_kernel<2, false><<<grid, NTH, 0, st>>>(A); break;
            }
        }
        SE3_LAUNCHED();
    }
    return SE3_OK;
}

extern "C" int se3_o3msg_edge_backward(const se3_o3msg_io* io, int32_t nio, int64_t n_dst, int64_t n_all,
                                       const int64_t* rowptr, const int64_t* tptr, const int32_t* perm, const float* y,
                                       int32_t ldy, const float* extra, int32_t ldx, const float* gpre, int32_t ldo,
                                       float* gdst, float* gsrc, int32_t ldt, float* gex, int32_t ldg, void* stream) {
    if (check_io(io, nio) || n_dst < 0 || n_all < 0 || ldt < 1 || ldy < 1 || ldo < 1) {
        set_error("o3msg backward: bad descriptor");
        return SE3_ERR_INVALID;
    }
    if ((n_dst > 0 && (!rowptr || !gdst)) || (n_all > 0 && (!tptr || !gsrc)) || ((n_dst > 0 || n_all > 0) && !y)) {
        set_error("o3msg backward: null argument");
        return SE3_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    for (int k = 0; k < nio; ++k) {
        if (io[k].nx > 0 && n_dst > 0 && (!extra || !gex || ldg < 1)) {
            set_error("o3msg backward: extras paths need extra and gex");
            return SE3_ERR_INVALID;
        }
        NodeArgs A;
        A.io = io[k]; A.y = y; A.ldy = ldy; A.ex = extra; A.ldx = ldx; A.gpre = gpre; A.ldo = ldo; A.ldt = ldt;
        A.gex = gex; A.ldg = ldg;
        A.per_block = NTH / io[k].mul;
        A.magic = (65536u + io[k].mul - 1) / io[k].mul;
        for (int role = 0; role < 2; ++role) {
            A.n = role ? n_all : n_dst;
            if (A.n == 0) continue;
            A.ptr = reinterpret_cast<const long long*>(role ? tptr : rowptr);
            A.perm = role ? perm : nullptr;
            A.g = role ? gsrc : gdst;
            const long long blocks = (A.n + A.per_block - 1) / A.per_block;
            if (blocks > 0x7fffffffLL) { set_error("o3msg backward: too many nodes"); return SE3_ERR_TOO_LARGE; }
            const int grid = (int)blocks;
            A.ny = sh_width(io, nio, ldy);
#define O3M_LAUNCH(LO_, SRC_)                                                              \
    do {                                                                                   \
        if (A.ny) o3msg_edge_bwd_kernel<LO_, SRC_, true><<<grid, NTH, 0, st>>>(A);          \
        else o3msg_edge_bwd_kernel<LO_, SRC_, false><<<grid, NTH, 0, st>>>(A);              \
    } while (0)
            if (role == 0) {
                switch (io[k].l) {
                    case 0: O3M_LAUNCH(0, false); break;
                    case 1: O3M_LAUNCH(1, false); break;
                    default: O3M_LAUNCH(2, false); break;
                }
            } else {
                switch (io[k].l) {
                    case 0: O3M_LAUNCH(0, true); break;
                    case 1: O3M_LAUNCH(1, true); break;
                    default: O3M_LAUNCH(2, true); break;
                }
            }
#undef O3M_LAUNCH
            SE3_LAUNCHED();
        }
    }
    return SE3_OK;
}
