// Gated non-linearity over flat irreps rows (public SEGNN: O3TensorProductSwishGate = TP -> Gate(scalars: silu, gates:
// sigmoid), both wrapped in normalize2mom), as a stand-alone elementwise kernel pair for layouts the fused epilogue of the
// l <= 1 tensor-product kernels does not cover (l = 2 blocks).  HBM-bound: every element is read / written once.
//   raw  = [ ns scalars | ng gate scalars | block 0: cnt0 x dim0 | block 1: cnt1 x dim1 | ... ],  ng = sum cnt
//   out  = [ cs silu(s) | block b, channel k, component c:  raw * cg sigmoid(gate[k_global]) ]
#include "common.cuh"

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

__global__ void __launch_bounds__(GATE_NT) gate_bwd_kernel(const __grid_constant__ GateL L, long long rows, int chunk,
                                                           unsigned magic, const float* __restrict__ raw,
                                                           const float* __restrict__ gout, const int32_t* __restrict__ gidx,
                                                           float* __restrict__ graw) {
    __shared__ GateTab T;
    gate_tables(L, T, true);
    const int d = L.d_raw;
    for (long long r0 = (long long)blockIdx.x * chunk; r0 < rows; r0 += (long long)gridDim.x * chunk) {
        const int nr = (int)min((long long)chunk, rows - r0);
        const float* x0 = raw + r0 * d;
        const float* g0 = gout + r0 * L.d_out;
        float* o0 = graw + r0 * d;
        for (int i = threadIdx.x; i < nr * d; i += GATE_NT) {
            const int r = (int)__umulhi((unsigned)i, magic), j = i - r * d;
            const float* x = x0 + r * d;
            const float* g = gidx ? gout + (long long)__ldg(gidx + r0 + r) * L.d_out : g0 + r * L.d_out;
            const int a = T.a[j], b = T.b[j], c = T.c[j];
            float v;
            if (c == 0) {                       // scalar: silu'
                const float s = x[j], sg = sigmoidf(s);
                v = g[j] * L.cs * (sg + s * sg * (1.f - sg));
            } else if (c < 0) {                 // gated value
                v = g[a] * (L.cg * sigmoidf(x[b]));
            } else {                            // gate scalar: sum over the components of its channel
                float acc = 0.f;
                for (int k = 0; k < c; ++k) acc = fmaf(g[b + k], x[a + k], acc);
                const float sg = sigmoidf(x[j]);
                v = acc * L.cg * sg * (1.f - sg);
            }
            o0[i] = v;
        }
    }
}

// out[n][:] = sum over the CSR row of node n of gate(raw[e][:]): the aggregation of the gated messages without the gated
// [E, d_out] tensor and without atomics (one warp per node, lane = output column, run-to-run deterministic).
__global__ void __launch_bounds__(GATE_NT) gate_segsum_kernel(const __grid_constant__ GateL L, long long n_seg,
                                                              const long long* __restrict__ rowptr,
                                                              const float* __restrict__ raw, float* __restrict__ out) {
    __shared__ GateTab T;
    gate_tables(L, T, false);
    const int lane = threadIdx.x & 31;
    const long long nwarp = (long long)gridDim.x * (GATE_NT / 32);
    for (long long n = (long long)blockIdx.x * (GATE_NT / 32) + (threadIdx.x >> 5); n < n_seg; n += nwarp) {
        const long long beg = __ldg(rowptr + n), end = __ldg(rowptr + n + 1);
        for (int c0 = 0; c0 < L.d_out; c0 += 128) {
            int aj[4], bj[4];
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int j = c0 + lane + 32 * k;
                aj[k] = j < L.d_out ? T.a[j] : -1;
                bj[k] = j < L.d_out ? T.b[j] : -1;
            }
#pragma unroll 2
            for (long long e = beg; e < end; ++e) {
                const float* x = raw + e * L.d_raw;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (aj[k] >= 0) {
                        const float v = __ldg(x + aj[k]);
                        acc[k] += bj[k] < 0 ? L.cs * v * sigmoidf(v) : v * (L.cg * sigmoidf(__ldg(x + bj[k])));
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (aj[k] >= 0) out[n * L.d_out + c0 + lane + 32 * k] = acc[k];
        }
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
    gate_chunk(L.d_raw, &chunk, &magic);
    const int grid = (int)std::min<long long>((rows + chunk - 1) / chunk, (long long)se3::num_sms() * 8);
    gate_bwd_kernel<<<grid, GATE_NT, 0, (cudaStream_t)stream>>>(L, rows, chunk, magic, raw, gout, nullptr, graw);
    SE3_LAUNCHED();
    return SE3_OK;
}

extern "C" int se3_gate_segment_sum_forward(int64_t n_seg, const int64_t* rowptr, int32_t ns, int32_t nblk, const int32_t* cnt,
                                            const int32_t* dim, float cs, float cg, const float* raw, float* out, void* stream) {
    GateL L;
    if (make_layout(L, ns, nblk, cnt, dim, cs, cg) || n_seg < 0 || (n_seg > 0 && (!rowptr || !out))) {
        se3::set_error("se3_gate_segment_sum_forward: bad argument");
        return SE3_ERR_INVALID;
    }
    if (n_seg == 0) return SE3_OK;
    const int grid = (int)std::min<long long>((n_seg + GATE_NT / 32 - 1) / (GATE_NT / 32), (long long)se3::num_sms() * 32);
    gate_segsum_kernel<<<grid, GATE_NT, 0, (cudaStream_t)stream>>>(L, n_seg, reinterpret_cast<const long long*>(rowptr), raw, out);
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
    gate_chunk(L.d_raw, &chunk, &magic);
    const int grid = (int)std::min<long long>((rows + chunk - 1) / chunk, (long long)se3::num_sms() * 8);
    gate_bwd_kernel<<<grid, GATE_NT, 0, (cudaStream_t)stream>>>(L, rows, chunk, magic, raw, gout, seg, graw);
    SE3_LAUNCHED();
    return SE3_OK;
}
