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

__device__ __forceinline__ float sigmoidf(float z) { return 1.0f / (1.0f + expf(-z)); }

// gate index and nothing else: column j (>= 0) of the gated part
__device__ __forceinline__ int gate_of(const GateL& L, int j) {
    int g0 = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        if (b < L.nblk) {
            const int w = L.cnt[b] * L.dim[b];
            if (j < L.off[b] + w) return g0 + (j - L.off[b]) / L.dim[b];
            g0 += L.cnt[b];
        }
    }
    return 0;
}

__global__ void gate_fwd_kernel(GateL L, long long rows, const float* __restrict__ raw, float* __restrict__ out) {
    const long long total = rows * L.d_out;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / L.d_out;
        const int j = (int)(i - r * L.d_out);
        const float* x = raw + r * L.d_raw;
        float v;
        if (j < L.ns) {
            const float s = x[j];
            v = L.cs * s * sigmoidf(s);
        } else {
            const int jj = j - L.ns;
            v = x[L.ns + L.ng + jj] * (L.cg * sigmoidf(x[L.ns + gate_of(L, jj)]));
        }
        out[i] = v;
    }
}

__global__ void gate_bwd_kernel(GateL L, long long rows, const float* __restrict__ raw, const float* __restrict__ gout,
                                float* __restrict__ graw) {
    const long long total = rows * L.d_raw;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / L.d_raw;
        const int j = (int)(i - r * L.d_raw);
        const float* x = raw + r * L.d_raw;
        const float* g = gout + r * L.d_out;
        float v;
        if (j < L.ns) {
            const float s = x[j], sg = sigmoidf(s);
            v = g[j] * L.cs * (sg + s * sg * (1.f - sg));
        } else if (j < L.ns + L.ng) {
            // gate scalar k: sum over the components of its channel
            int k = j - L.ns, b = 0, g0 = 0;
            while (b + 1 < L.nblk && k >= g0 + L.cnt[b]) { g0 += L.cnt[b]; ++b; }
            const int col = L.off[b] + (k - g0) * L.dim[b];
            float acc = 0.f;
            for (int c = 0; c < L.dim[b]; ++c) acc += g[L.ns + col + c] * x[L.ns + L.ng + col + c];
            const float sg = sigmoidf(x[j]);
            v = acc * L.cg * sg * (1.f - sg);
        } else {
            const int jj = j - L.ns - L.ng;
            v = g[L.ns + jj] * (L.cg * sigmoidf(x[L.ns + gate_of(L, jj)]));
        }
        graw[i] = v;
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
    return L.d_out > 0 ? SE3_OK : SE3_ERR_INVALID;
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
    const long long total = rows * L.d_out;
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)se3::num_sms() * 16);
    gate_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(L, rows, raw, out);
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
    const long long total = rows * L.d_raw;
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)se3::num_sms() * 16);
    gate_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(L, rows, raw, gout, graw);
    SE3_LAUNCHED();
    return SE3_OK;
}
