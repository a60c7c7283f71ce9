// Node-level weight contraction of message 1 (the dense GEMM over irrep channels that csrc/msg_table.cu's header
// describes), block-sparse and in exact fp32:
//     T[n][half][ch][0]   = sum_k  x[n][k]            Ws[half][k][ch]      (k < NS scalars)
//     T[n][half][ch][1+c] = sum_k  x[n][NS + 3k + c]  Wv[half][k][ch]      (k < NV vectors, c = x,y,z)
// with Ws / Wv = the rows of weights_l0e | weights_l1o that belong to the dst / src half of the concatenated input
// (L1TP:81-88 row order), norms and 1/sqrt(3) folded in.  As a dense [D x 8 CH] matrix three quarters of the entries
// are structural zeros (a scalar only feeds the P column, a vector component only its own U column), which is why
// these are hand-written kernels and not a library GEMM: 4x fewer multiply-adds, no expanded weight matrix, and the
// weight gradient comes out directly in the parameters' layout.
//   msg1_node_table   T = x . W                      (forward)
//   msg1_node_gx      gx = G . W^T                   (input gradient)
//   msg1_node_gw      gW partials = x^T . G  per block, msg1_node_gw_reduce -> gwz / gwv (deterministic)
// All three: persistent blocks, weights resident in shared memory, node tiles staged in shared memory, register
// accumulators; HBM traffic = every operand once.
#include <algorithm>

#include "tc_common.cuh"

namespace se3 {

template <int NS, int NV>
struct NodeDims {
    static constexpr int MZ = NS + NV, CH = NS + 2 * NV, D = NS + 3 * NV, HALF = 4 * CH, LDT = 2 * HALF;
    static constexpr int KR = NS + NV;              // weight rows per half (scalars, then vectors)
    static constexpr int WF = 2 * KR * CH;          // folded weights: [half][KR][CH]
    static constexpr int NSC = 2 * NS + 2;          // scalar rows of the parameter (dst, src, extras)
};

// folded weight of (half p, row k (scalars first), channel ch) from the parameters
template <int NS, int NV>
__device__ __forceinline__ float node_w(const float* wz, const float* wv, const float* nz, const float* nvn, int p, int k, int ch) {
    using Nd = NodeDims<NS, NV>;
    const bool vec = k >= NS;
    const int row = vec ? Nd::NSC + p * NV + (k - NS) : p * NS + k;
    if (ch < Nd::MZ) return (vec ? C3f : 1.0f) * (nz ? __ldg(nz + ch) : 1.0f) * __ldg(wz + row * Nd::MZ + ch);
    return C3f * (nvn ? __ldg(nvn + 3 * (ch - Nd::MZ)) : 1.0f) * __ldg(wv + row * NV + ch - Nd::MZ);
}

static constexpr int NT_TILE = 8;      // nodes per thread (register blocking)

// ---------------------------------------------------------------- forward: T = x . W
template <int NS, int NV>
__global__ void __launch_bounds__(128) msg1_node_table_kernel(const float* __restrict__ x, long long n, const float* wz,
                                                               const float* wv, const float* nz, const float* nvn,
                                                               float* __restrict__ T) {
    using Nd = NodeDims<NS, NV>;
    constexpr int CH = Nd::CH, D = Nd::D, KR = Nd::KR;
    static_assert(2 * CH <= 128, "one thread per (half, channel)");
    __shared__ float ws[Nd::WF];
    __shared__ __align__(16) float xs[NT_TILE][D];
    const int tid = threadIdx.x;
    for (int t = tid; t < Nd::WF; t += 128) {
        const int p = t / (KR * CH), r = t - p * KR * CH, k = r / CH, ch = r - k * CH;
        ws[t] = node_w<NS, NV>(wz, wv, nz, nvn, p, k, ch);
    }
    const bool act = tid < 2 * CH;
    const int p = act ? tid / CH : 0, ch = act ? tid - p * CH : 0;
    const float* wp = ws + p * KR * CH + ch;
    const long long ntile = (n + NT_TILE - 1) / NT_TILE;
    for (long long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const long long n0 = tile * NT_TILE;
        __syncthreads();
        for (int t = tid; t < NT_TILE * D; t += 128) {
            const long long row = n0 + t / D;
            (&xs[0][0])[t] = row < n ? __ldg(x + n0 * D + t) : 0.0f;
        }
        __syncthreads();
        if (!act) continue;
        float acc[NT_TILE][4];
#pragma unroll
        for (int r = 0; r < NT_TILE; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.0f;
#pragma unroll 2
        for (int k = 0; k < NS; ++k) {
            const float w = wp[k * CH];
#pragma unroll
            for (int r = 0; r < NT_TILE; ++r) acc[r][0] = fmaf(xs[r][k], w, acc[r][0]);
        }
#pragma unroll 2
        for (int k = 0; k < NV; ++k) {
            const float w = wp[(NS + k) * CH];
#pragma unroll
            for (int r = 0; r < NT_TILE; ++r) {
                acc[r][1] = fmaf(xs[r][NS + 3 * k], w, acc[r][1]);
                acc[r][2] = fmaf(xs[r][NS + 3 * k + 1], w, acc[r][2]);
                acc[r][3] = fmaf(xs[r][NS + 3 * k + 2], w, acc[r][3]);
            }
        }
#pragma unroll
        for (int r = 0; r < NT_TILE; ++r)
            if (n0 + r < n)
                *reinterpret_cast<float4*>(T + (n0 + r) * Nd::LDT + p * Nd::HALF + 4 * ch) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    }
}

// ---------------------------------------------------------------- input gradient: gx = G . W^T
template <int NS, int NV>
__global__ void __launch_bounds__(128) msg1_node_gx_kernel(const float* __restrict__ G, long long n, const float* wz,
                                                            const float* wv, const float* nz, const float* nvn,
                                                            float* __restrict__ gx) {
    using Nd = NodeDims<NS, NV>;
    constexpr int CH = Nd::CH, D = Nd::D, KR = Nd::KR, LDT = Nd::LDT;
    static_assert(D <= 64, "one thread per feature column, two node groups per block");
    // transposed folded weights: wt[half][ch][KR] (threads of a warp differ in k: conflict-free)
    extern __shared__ __align__(16) float smn[];
    float* wt = smn;                                  // WF
    float* gs = smn + Nd::WF;                         // [2 NT_TILE][LDT]
    const int tid = threadIdx.x;
    for (int t = tid; t < Nd::WF; t += 128) {
        const int p = t / (KR * CH), r = t - p * KR * CH, c = r / KR, k = r - c * KR;
        wt[t] = node_w<NS, NV>(wz, wv, nz, nvn, p, k, c);
    }
    const int col = tid & 63, grp = tid >> 6;
    const bool act = col < D;
    // column -> (weight row k, component j of the table entry)
    const int k = col < NS ? col : NS + (col - NS) / 3;
    const int j = col < NS ? 0 : 1 + (col - NS) % 3;
    const long long ntile = (n + 2 * NT_TILE - 1) / (2 * NT_TILE);
    for (long long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const long long n0 = tile * 2 * NT_TILE;
        __syncthreads();
        for (int t = tid; t < 2 * NT_TILE * LDT / 4; t += 128) {
            const long long row = n0 + (t * 4) / LDT;
            reinterpret_cast<float4*>(gs)[t] = row < n ? __ldg(reinterpret_cast<const float4*>(G + n0 * LDT) + t) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
        if (!act) continue;
        float acc[NT_TILE];
#pragma unroll
        for (int r = 0; r < NT_TILE; ++r) acc[r] = 0.0f;
        const float* gr = gs + grp * NT_TILE * LDT + j;
#pragma unroll 2
        for (int pc = 0; pc < 2 * CH; ++pc) {          // pc = half * CH + ch
            const float w = wt[pc * KR + k];
#pragma unroll
            for (int r = 0; r < NT_TILE; ++r) acc[r] = fmaf(gr[r * LDT + 4 * pc], w, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < NT_TILE; ++r) {
            const long long row = n0 + grp * NT_TILE + r;
            if (row < n) gx[row * D + col] = acc[r];
        }
    }
}

// ---------------------------------------------------------------- weight gradient: per-block partials of x^T . G
// thread = (half, channel, row group): the scalar rows are split into three groups, the vector rows form the fourth
template <int NS, int NV>
struct GwMap {
    static constexpr int G0 = (NS + 2) / 3, G1 = (NS - G0 + 1) / 2, G2 = NS - G0 - G1;      // scalar rows per group
    static constexpr int MAXA = (G0 > NV ? G0 : NV);
};
static constexpr int GW_TILE = 16;     // nodes staged per step

template <int NS, int NV>
__global__ void __launch_bounds__(448) msg1_node_gw_kernel(const float* __restrict__ x, const float* __restrict__ G, long long n,
                                                            float* __restrict__ part) {
    using Nd = NodeDims<NS, NV>;
    using Gm = GwMap<NS, NV>;
    constexpr int CH = Nd::CH, D = Nd::D, KR = Nd::KR, LDT = Nd::LDT;
    static_assert(8 * CH <= 448, "one thread per (half, channel, row group)");
    extern __shared__ __align__(16) float smn[];
    float* xs = smn;                                  // [GW_TILE][D]
    float* gs = smn + GW_TILE * D;                    // [GW_TILE][LDT]
    const int tid = threadIdx.x;
    const bool act = tid < 8 * CH;
    const int grp = act ? tid / (2 * CH) : 0, pc = act ? tid - grp * 2 * CH : 0;      // pc = half * CH + ch
    const int kbeg = grp == 0 ? 0 : (grp == 1 ? Gm::G0 : (grp == 2 ? Gm::G0 + Gm::G1 : NS));
    const int kcnt = grp == 0 ? Gm::G0 : (grp == 1 ? Gm::G1 : (grp == 2 ? Gm::G2 : NV));
    float acc[Gm::MAXA];
#pragma unroll
    for (int i = 0; i < Gm::MAXA; ++i) acc[i] = 0.0f;
    const long long ntile = (n + GW_TILE - 1) / GW_TILE;
    for (long long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const long long n0 = tile * GW_TILE;
        __syncthreads();
        for (int t = tid; t < GW_TILE * D; t += 448) xs[t] = n0 + t / D < n ? __ldg(x + n0 * D + t) : 0.0f;
        for (int t = tid; t < GW_TILE * LDT / 4; t += 448) {
            const long long row = n0 + (t * 4) / LDT;
            reinterpret_cast<float4*>(gs)[t] = row < n ? __ldg(reinterpret_cast<const float4*>(G + n0 * LDT) + t) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
        if (!act) continue;
        if (grp < 3) {
#pragma unroll 4
            for (int r = 0; r < GW_TILE; ++r) {
                const float g = reinterpret_cast<const float4*>(gs + r * LDT + 4 * pc)->x;   // 16-byte load: conflict-free
                const float* xr = xs + r * D + kbeg;
#pragma unroll
                for (int i = 0; i < Gm::MAXA; ++i)
                    if (i < kcnt) acc[i] = fmaf(xr[i], g, acc[i]);
            }
        } else {
#pragma unroll 4
            for (int r = 0; r < GW_TILE; ++r) {
                const float4 g = *reinterpret_cast<const float4*>(gs + r * LDT + 4 * pc);
                const float* xr = xs + r * D + NS;
#pragma unroll
                for (int i = 0; i < NV; ++i) acc[i] = fmaf(xr[3 * i], g.y, fmaf(xr[3 * i + 1], g.z, fmaf(xr[3 * i + 2], g.w, acc[i])));
            }
        }
    }
    if (act) {
        float* po = part + (long long)blockIdx.x * Nd::WF + (pc / CH) * KR * CH + (pc % CH);   // [half][k][ch]
#pragma unroll
        for (int i = 0; i < Gm::MAXA; ++i)
            if (i < kcnt) po[(kbeg + i) * CH] = acc[i];
    }
}

// gwz [(2NS+2+2NV), MZ], gwv [(same), NV] overwritten: sum of the per-block partials (fixed order) with the folded
// factors, plus the extras' rows from the per-block partials of the edge kernel
template <int NS, int NV>
__global__ void __launch_bounds__(256) msg1_node_gw_reduce_kernel(const float* __restrict__ part, int nparts,
                                                                  const float* __restrict__ gwe_part, int neparts, const float* nz,
                                                                  const float* nvn, float* __restrict__ gwz, float* __restrict__ gwv) {
    using Nd = NodeDims<NS, NV>;
    constexpr int MZ = Nd::MZ, CH = Nd::CH, KR = Nd::KR, NSC = Nd::NSC, ROWS = NSC + 2 * NV;
    // one warp per output element: lane l sums partials l, l + 32, ... (fixed order), then a shuffle tree
    const int lane = threadIdx.x & 31;
    for (int t = blockIdx.x * 8 + (threadIdx.x >> 5); t < ROWS * CH; t += gridDim.x * 8) {
        const int row = t / CH, ch = t - row * CH;
        float g = 0.0f;
        bool vec = false;
        if (row >= 2 * NS && row < NSC) {
            const int j = row - 2 * NS;
            for (int q = lane; q < neparts; q += 32) g += gwe_part[((long long)q * 2 + j) * CH + ch];
        } else {
            int p, k;
            if (row < 2 * NS) { p = row / NS; k = row - p * NS; }
            else { const int r = row - NSC; p = r / NV; k = NS + r - p * NV; vec = true; }
            const float* src = part + (p * KR + k) * CH + ch;
            for (int q = lane; q < nparts; q += 32) g += src[(long long)q * Nd::WF];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
        if (lane == 0) {
            if (ch < MZ) gwz[row * MZ + ch] = g * (vec ? C3f : 1.0f) * (nz ? nz[ch] : 1.0f);
            else gwv[row * NV + ch - MZ] = g * C3f * (nvn ? nvn[3 * (ch - MZ)] : 1.0f);
        }
    }
}

template <int NS, int NV>
static int node_table(const float* x, long long n, const float* wz, const float* wv, const float* nz, const float* nvn, float* T,
                      cudaStream_t st) {
    const int grid = (int)std::max<long long>(1, std::min<long long>((n + NT_TILE - 1) / NT_TILE, (long long)num_sms() * 8));
    msg1_node_table_kernel<NS, NV><<<grid, 128, 0, st>>>(x, n, wz, wv, nz, nvn, T);
    SE3_LAUNCHED();
    return SE3_OK;
}

template <int NS, int NV>
static int node_backward(const float* x, const float* G, long long n, const float* wz, const float* wv, const float* nz,
                         const float* nvn, const float* gwe_part, int neparts, float* gx, float* gwz, float* gwv, float* part,
                         int max_parts, cudaStream_t st) {
    using Nd = NodeDims<NS, NV>;
    if (gx) {
        const int smem = (Nd::WF + 2 * NT_TILE * Nd::LDT) * 4;
        static bool attr = false;
        if (!attr) { SE3_CUDA_TRY(cudaFuncSetAttribute(msg1_node_gx_kernel<NS, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr = true; }
        const int grid = (int)std::max<long long>(1, std::min<long long>((n + 2 * NT_TILE - 1) / (2 * NT_TILE), (long long)num_sms() * 4));
        msg1_node_gx_kernel<NS, NV><<<grid, 128, smem, st>>>(G, n, wz, wv, nz, nvn, gx);
        SE3_LAUNCHED();
    }
    if (gwz && gwv) {
        const int smem = GW_TILE * (Nd::D + Nd::LDT) * 4;
        static bool attr = false;
        if (!attr) { SE3_CUDA_TRY(cudaFuncSetAttribute(msg1_node_gw_kernel<NS, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr = true; }
        const int grid = (int)std::max<long long>(1, std::min<long long>((n + GW_TILE - 1) / GW_TILE, std::min<long long>(max_parts, (long long)num_sms() * 4)));
        msg1_node_gw_kernel<NS, NV><<<grid, 448, smem, st>>>(x, G, n, part);
        SE3_LAUNCHED();
        msg1_node_gw_reduce_kernel<NS, NV><<<num_sms() * 2, 256, 0, st>>>(part, grid, gwe_part, neparts, nz, nvn, gwz, gwv);
        SE3_LAUNCHED();
    }
    return SE3_OK;
}

}  // namespace se3

using namespace se3;

#define SE3_NODE_DISPATCH(ns, nv, CALL)                                      \
    if ((ns) == 34 && (nv) == 10) { CALL(34, 10) }                           \
    else if ((ns) == 16 && (nv) == 8) { CALL(16, 8) }                        \
    else if ((ns) == 8 && (nv) == 4) { CALL(8, 4) }                          \
    else { set_error("msg1 node kernels: hidden irreps %dx0e+%dx1o are not instantiated", (int)(ns), (int)(nv)); return SE3_ERR_INVALID; }

extern "C" int se3_msg1_node_parts(int32_t ns, int32_t nv, int32_t* max_parts, int32_t* part_floats) {
    if (!max_parts || !part_floats) { set_error("msg1_node_parts: null argument"); return SE3_ERR_INVALID; }
    *max_parts = num_sms() * 4;
    *part_floats = 2 * (ns + nv) * (ns + 2 * nv);
    return SE3_OK;
}

extern "C" int se3_msg1_node_table(int32_t ns, int32_t nv, int64_t n, const float* x, const float* wz, const float* wv,
                                   const float* nz, const float* nvn, float* table, void* stream) {
    if (n < 0 || (n > 0 && (!x || !wz || !wv || !table))) { set_error("msg1_node_table: bad argument"); return SE3_ERR_INVALID; }
    if (n == 0) return SE3_OK;
#define CALL(a, b) return node_table<a, b>(x, n, wz, wv, nz, nvn, table, (cudaStream_t)stream);
    SE3_NODE_DISPATCH(ns, nv, CALL)
#undef CALL
}

extern "C" int se3_msg1_node_backward(int32_t ns, int32_t nv, int64_t n, const float* x, const float* G, const float* wz,
                                      const float* wv, const float* nz, const float* nvn, const float* gwe_part,
                                      int32_t neparts, float* gx, float* gwz, float* gwv, float* part, int32_t max_parts,
                                      void* stream) {
    if (n <= 0 || !x || !G || !wz || !wv || (gwz && (!gwv || !part || !gwe_part || max_parts < 1))) {
        set_error("msg1_node_backward: bad argument");
        return SE3_ERR_INVALID;
    }
#define CALL(a, b) return node_backward<a, b>(x, G, n, wz, wv, nz, nvn, gwe_part, neparts, gx, gwz, gwv, part, max_parts, (cudaStream_t)stream);
    SE3_NODE_DISPATCH(ns, nv, CALL)
#undef CALL
}
