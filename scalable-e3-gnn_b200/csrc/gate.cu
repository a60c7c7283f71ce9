// Gated non-linearity over flat irreps rows (public SEGNN: O3TensorProductSwishGate = TP -> Gate(scalars: silu, gates:
// sigmoid), both wrapped in normalize2mom), as a stand-alone elementwise kernel pair for layouts the fused epilogue of the
// l <= 1 tensor-product kernels does not cover (l = 2 blocks).  HBM-bound: every element is read / written once.
//   raw  = [ ns scalars | ng gate scalars | block 0: cnt0 x dim0 | block 1: cnt1 x dim1 | ... ],  ng = sum cnt
//   out  = [ cs silu(s) | block b, channel k, component c:  raw * cg sigmoid(gate[k_global]) ]
#include "tc_common.cuh"

namespace {

struct GateL {
    int ns, ng, nblk, d_raw, d_out;
    int cnt[4], dim[4], off[4];  // off: first column of the block relative to the start of the gated part
    float cs, cg;
};

// sigmoid with two MUFU ops (ex2.approx, rcp.approx: ~1e-7 relative, far inside the 1e-5 parity budget)
__device__ __forceinline__ float sigmoidf(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}

constexpr int GATE_NT = 256;
constexpr int GATE_MAXD = 512;   // widest raw row the column tables in shared memory cover

// Column tables (shared memory, built once per block from the layout):
//   forward, output column j:   src[j] = raw column of the value, gate[j] = raw column of its gate scalar (-1: silu)
//   backward, raw column j:     scalar: (j, -1, 0) | gate scalar: (first gated raw column, first gout column, dim) |
//                               gated value: (gout column, gate raw column, -1)
// A block walks chunks of whole rows, thread = element, so global reads and writes are coalesced and the only
// division is a 32-bit multiply-high by a per-launch constant.
struct GateTab {
    short a[GATE_MAXD], b[GATE_MAXD], c[GATE_MAXD];
};

__device__ __forceinline__ void gate_tables(const GateL& L, GateTab& T, bool bwd) {
    for (int j = threadIdx.x; j < (bwd ? L.d_raw : L.d_out); j += blockDim.x) {
        short a = (short)j, b = -1, c = 0;
        if (!bwd) {
            if (j >= L.ns) {
                int jj = j - L.ns, g0 = 0;
                a = (short)(L.ns + L.ng + jj);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < L.nblk) {
                        if (jj >= L.off[k] && jj < L.off[k] + L.cnt[k] * L.dim[k]) b = (short)(L.ns + g0 + (jj - L.off[k]) / L.dim[k]);
                        g0 += L.cnt[k];
                    }
            }
        } else if (j >= L.ns && j < L.ns + L.ng) {
            int k0 = j - L.ns, g0 = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < L.nblk) {
                    if (k0 >= g0 && k0 < g0 + L.cnt[k]) {
                        const int col = L.off[k] + (k0 - g0) * L.dim[k];
                        a = (short)(L.ns + L.ng + col); b = (short)(L.ns + col); c = (short)L.dim[k];
                    }
                    g0 += L.cnt[k];
                }
        } else if (j >= L.ns + L.ng) {
            int jj = j - L.ns - L.ng, g0 = 0;
            a = (short)(L.ns + jj); c = -1;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < L.nblk) {
                    if (jj >= L.off[k] && jj < L.off[k] + L.cnt[k] * L.dim[k]) b = (short)(L.ns + g0 + (jj - L.off[k]) / L.dim[k]);
                    g0 += L.cnt[k];
                }
        }
        T.a[j] = a; T.b[j] = b; T.c[j] = c;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(GATE_NT) gate_fwd_kernel(const __grid_constant__ GateL L, long long rows, int chunk,
                                                           unsigned magic, const float* __restrict__ raw,
                                                           float* __restrict__ out) {
    __shared__ GateTab T;
    gate_tables(L, T, false);
    const int d = L.d_out;
    for (long long r0 = (long long)blockIdx.x * chunk; r0 < rows; r0 += (long long)gridDim.x * chunk) {
        const int nr = (int)min((long long)chunk, rows - r0);
        const float* x0 = raw + r0 * L.d_raw;
        float* o0 = out + r0 * d;
        for (int i = threadIdx.x; i < nr * d; i += GATE_NT) {
            const int r = (int)__umulhi((unsigned)i, magic), j = i - r * d;
            const float* x = x0 + r * L.d_raw;
            const float v = x[T.a[j]];
            const int gc = T.b[j];
            o0[i] = gc < 0 ? L.cs * v * sigmoidf(v) : v * (L.cg * sigmoidf(x[gc]));
        }
    }
}

// Backward, thread = (row, unit): a unit is one scalar or one gated channel with all its components, so the gate's
// sigmoid is evaluated once per channel and the reduction over the components stays in registers (the element-wise
// version spent 130 instructions per element, a fifth of them on the row pointers).  gidx: optional row index of the
// cotangent (the fused gate + segment sum reads the cotangent of the destination node).
__global__ void __launch_bounds__(GATE_NT) gate_bwd_kernel(const __grid_constant__ GateL L, long long rows, int chunk,
                                                           unsigned magic, const float* __restrict__ raw,
                                                           const float* __restrict__ gout, const int32_t* __restrict__ gidx,
                                                           float* __restrict__ graw) {
    __shared__ short ccol[GATE_MAXD], cdim[GATE_MAXD];
    for (int k = threadIdx.x; k < L.ng; k += blockDim.x) {
        int g0 = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if (b < L.nblk) {
                if (k >= g0 && k < g0 + L.cnt[b]) { ccol[k] = (short)(L.off[b] + (k - g0) * L.dim[b]); cdim[k] = (short)L.dim[b]; }
                g0 += L.cnt[b];
            }
    }
    __syncthreads();
    const int nu = L.ns + L.ng, d = L.d_raw, gbase = L.ns + L.ng;
    for (long long r0 = (long long)blockIdx.x * chunk; r0 < rows; r0 += (long long)gridDim.x * chunk) {
        const int nr = (int)min((long long)chunk, rows - r0);
        const float* x0 = raw + r0 * d;
        float* o0 = graw + r0 * d;
        for (int i = threadIdx.x; i < nr * nu; i += GATE_NT) {
            const int r = (int)__umulhi((unsigned)i, magic), u = i - r * nu;
            const float* x = x0 + r * d;
            float* o = o0 + r * d;
            const float* g = gout + (gidx ? (long long)__ldg(gidx + r0 + r) : r0 + r) * L.d_out;
            if (u < L.ns) {                       // scalar: silu'
                const float sv = x[u], sg = sigmoidf(sv);
                o[u] = g[u] * L.cs * (sg + sv * sg * (1.f - sg));
            } else {                              // gated channel: its components and its gate scalar
                const int k = u - L.ns, col = ccol[k], dim = cdim[k];
                const float sg = sigmoidf(x[u]), f = L.cg * sg;
                const float* gv = g + L.ns + col;
                const float* xv = x + gbase + col;
                float* ov = o + gbase + col;
                float acc = 0.f;
                for (int c = 0; c < dim; ++c) {
                    const float gc = gv[c];
                    ov[c] = gc * f;
                    acc = fmaf(gc, xv[c], acc);
                }
                o[u] = acc * f * (1.f - sg);
            }
        }
    }
}

// out[seg[e]][:] += gate(raw[e][:]) for edges sorted by segment: the aggregation of the gated messages without the gated
// [E, d_out] tensor.  Tiles of 64 consecutive edges: the raw rows are one contiguous block (a linear 16-byte copy into
// shared memory), the gate is applied element-wise into a second tile, and the runs of equal segment are summed by
// sorted_segment_sum_tile (tc_common.cuh): plain stores for runs inside the tile, atomic adds only for the (at most two)
// runs shared with the neighbouring tiles; out is zeroed by the host function.
constexpr int GS_TM = 64;
__global__ void __launch_bounds__(GATE_NT) gate_segsum_kernel(const __grid_constant__ GateL L, long long rows,
                                                              const int32_t* __restrict__ seg,
                                                              const float* __restrict__ raw, float* __restrict__ out) {
    extern __shared__ __align__(16) float gs_sm[];
    __shared__ GateTab T;
    gate_tables(L, T, false);
    float* rawt = gs_sm;                                     // [64][d_raw], exact image of the global rows
    float* gt = rawt + ((GS_TM * L.d_raw + 3) & ~3);         // [64][d_out]
    float4* hs = reinterpret_cast<float4*>(gt + GS_TM * L.d_out);
    int* sseg = reinterpret_cast<int*>(hs + 8 * L.d_out);    // [66]
    const int tid = threadIdx.x;
    const long long ntiles = (rows + GS_TM - 1) / GS_TM;
    const bool v4 = (reinterpret_cast<uintptr_t>(raw) & 15) == 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long r0 = tile * GS_TM;
        const int nv = (int)min((long long)GS_TM, rows - r0), total = nv * L.d_raw;
        const float* src = raw + r0 * L.d_raw;               // r0 * d_raw * 4 is a multiple of 256 bytes
        if (v4) {
            for (int i = tid; i < (total >> 2); i += GATE_NT)
                reinterpret_cast<float4*>(rawt)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
            for (int i = (total & ~3) + tid; i < total; i += GATE_NT) rawt[i] = __ldg(src + i);
        } else {
            for (int i = tid; i < total; i += GATE_NT) rawt[i] = __ldg(src + i);
        }
        if (tid < nv) sseg[1 + tid] = __ldg(seg + r0 + tid);
        if (tid == GS_TM) sseg[0] = r0 > 0 ? __ldg(seg + r0 - 1) : -1;
        if (tid == GS_TM + 1) sseg[GS_TM + 1] = r0 + GS_TM < rows ? __ldg(seg + r0 + GS_TM) : -1;
        __syncthreads();
        for (int r = tid >> 5; r < nv; r += GATE_NT / 32) {  // warp = row, lane = output column
            const float* x = rawt + r * L.d_raw;
            for (int j = tid & 31; j < L.d_out; j += 32) {
                const float v = x[T.a[j]];
                const int gc = T.b[j];
                gt[r * L.d_out + j] = gc < 0 ? L.cs * v * sigmoidf(v) : v * (L.cg * sigmoidf(x[gc]));
            }
        }
        __syncthreads();
        se3::sorted_segment_sum_tile<GATE_NT>(gt, L.d_out, L.d_out, nv, sseg, out, L.d_out, hs, tid, 1);
        __syncthreads();
    }
}

int make_layout(GateL& L, int ns, int nblk, const int32_t* cnt, const int32_t* dim, float cs, float cg) {
    if (ns < 0 || nblk < 0 || nblk > 4 || (nblk > 0 && (!cnt || !dim))) return SE3_ERR_INVALID;
    L.ns = ns; L.nblk = nblk; L.cs = cs; L.cg = cg; L.ng = 0;
    int off = 0;
    for (int b = 0; b < 4; ++b) {
        L.cnt[b] = b < nblk ? cnt[b] : 0;
        L.dim[b] = b < nblk ? dim[b] : 1;
        if (b < nblk && (cnt[b] < 1 || dim[b] < 1)) return SE3_ERR_INVALID;
        L.off[b] = off;
        off += L.cnt[b] * L.dim[b];
        L.ng += L.cnt[b];
    }
    L.d_out = ns + off;
    L.d_raw = ns + L.ng + off;
    return L.d_out > 0 && L.d_raw <= GATE_MAXD ? SE3_OK : SE3_ERR_INVALID;
}

// rows per block iteration: ~16 elements per thread, whole rows; the element index inside a chunk stays far below 2^31,
// and floor(i / d) == umulhi(i, ceil(2^32 / d)) holds for i < 2^16 * ... (checked for the chunk sizes used: i < 8192)
void gate_chunk(int d, int* chunk, unsigned* magic) {
    *chunk = std::max(1, (16 * GATE_NT) / d);
    *magic = (unsigned)((0x100000000ull + (unsigned long long)d - 1) / (unsigned long long)d);
}

}  // namespace

extern "C" int se3_gate_forward(int64_t rows, int32_t ns, int32_t nblk, const int32_t* cnt, const int32_t* dim, float cs,
                                float cg, const float* raw, float* out, void* stream) {
    GateL L;
    if (make_layout(L, ns, nblk, cnt, dim, cs, cg) || rows < 0 || (rows > 0 && (!raw || !out))) {
        se3::set_error("se3_gate_forward: bad argument");
        return SE3_ERR_INVALID;
    }
    if (rows == 0) return SE3_OK;
    int chunk;
    unsigned magic;
    gate_chunk(L.d_out, &chunk, &magic);
    const int grid = (int)std::min<long long>((rows + chunk - 1) / chunk, (long long)se3::num_sms() * 8);
    gate_fwd_kernel<<<grid, GATE_NT, 0, (cudaStream_t)stream>>>(L, rows, chunk, magic, raw, out);
    SE3_LAUNCHED();
    return SE3_OK;
}

extern "C" int se3_gate_backward(int64_t rows, int32_t ns, int32_t nblk, const int32_t* cnt, const int32_t* dim, float cs,
                                 float cg, const float* raw, const float* gout, float* graw, void* stream) {
    GateL L;
    if (make_layout(L, ns, nblk, cnt, dim, cs, cg) || rows < 0 || (rows > 0 && (!raw || !gout || !graw))) {
        se3::set_error("se3_gate_backward: bad argument");
        return SE3_ERR_INVALID;
    }
    if (rows == 0) return SE3_OK;
    int chunk;
    unsigned magic;
    gate_chunk(L.ns + L.ng, &chunk, &magic);
    const int grid = (int)std::min<long long>((rows + chunk - 1) / chunk, (long long)se3::num_sms() * 8);
    gate_bwd_kernel<<<grid, GATE_NT, 0, (cudaStream_t)stream>>>(L, rows, chunk, magic, raw, gout, nullptr, graw);
    SE3_LAUNCHED();
    return SE3_OK;
}

extern "C" int se3_gate_segment_sum_forward(int64_t rows, const int32_t* seg, int64_t n_seg, int32_t ns, int32_t nblk,
                                            const int32_t* cnt, const int32_t* dim, float cs, float cg, const float* raw,
                                            float* out, void* stream) {
    GateL L;
    if (make_layout(L, ns, nblk, cnt, dim, cs, cg) || rows < 0 || n_seg < 0 || (n_seg > 0 && !out) || (rows > 0 && (!seg || !raw))) {
        se3::set_error("se3_gate_segment_sum_forward: bad argument");
        return SE3_ERR_INVALID;
    }
    if (n_seg == 0) return SE3_OK;
    SE3_CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)n_seg * L.d_out, (cudaStream_t)stream));
    if (rows == 0) return SE3_OK;
    const size_t smem = 4 * (((size_t)GS_TM * L.d_raw + 3) / 4 * 4 + (size_t)GS_TM * L.d_out + 4 * 8 * L.d_out + GS_TM + 2);
    if (smem > 200 * 1024) {
        se3::set_error("se3_gate_segment_sum_forward: rows too wide for the shared-memory tile");
        return SE3_ERR_TOO_LARGE;
    }
    static size_t attr = 48 * 1024;
    if (smem > attr) {
        SE3_CUDA_TRY(cudaFuncSetAttribute(gate_segsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    const long long ntiles = (rows + GS_TM - 1) / GS_TM;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(6, (200 * 1024) / smem));
    const int grid = (int)std::min<long long>(ntiles, (long long)se3::num_sms() * per_sm);
    gate_segsum_kernel<<<grid, GATE_NT, smem, (cudaStream_t)stream>>>(L, rows, seg, raw, out);
    SE3_LAUNCHED();
    return SE3_OK;
}

extern "C" int se3_gate_segment_sum_backward(int64_t rows, const int32_t* seg, int32_t ns, int32_t nblk, const int32_t* cnt,
                                             const int32_t* dim, float cs, float cg, const float* raw, const float* gout,
                                             float* graw, void* stream) {
    GateL L;
    if (make_layout(L, ns, nblk, cnt, dim, cs, cg) || rows < 0 || (rows > 0 && (!seg || !raw || !gout || !graw))) {
        se3::set_error("se3_gate_segment_sum_backward: bad argument");
        return SE3_ERR_INVALID;
    }
    if (rows == 0) return SE3_OK;
    int chunk;
    unsigned magic;
    gate_chunk(L.ns + L.ng, &chunk, &magic);
    const int grid = (int)std::min<long long>((rows + chunk - 1) / chunk, (long long)se3::num_sms() * 8);
    gate_bwd_kernel<<<grid, GATE_NT, 0, (cudaStream_t)stream>>>(L, rows, chunk, magic, raw, gout, seg, graw);
    SE3_LAUNCHED();
    return SE3_OK;
}
