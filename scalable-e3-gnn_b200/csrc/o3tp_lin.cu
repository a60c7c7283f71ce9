// l <= 2 tensor product with a SCALAR second input (in2 = one l = 0 irrep): every path is a plain linear map
//     out[n][io][w][i] = a_io / sqrt(2l+1) * y[n] * sum_u W_p[u][w] x[n][i1][u][i]        (l = l_i1 = l_io)
// This is the node-level contraction of the message product by linearity (se3gnn_b200/o3msg.py: the tables
// T = O3TP(x, 1; W_role) with 446 output columns per node); the general kernels of o3tp.cu need 148 KB of shared memory
// per CTA for that shape (one CTA per SM, 4.5 ms backward for 1.1M rows).  Here: one warp per row, the row staged in
// shared memory, the weights resident in shared memory, lane = output element; forward and input gradients.  (The
// weight gradient of these products runs on the tensor cores, o3tp_tc_gw.cu.)
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "o3tp_lin.h"
#include "common.cuh"

namespace {

constexpr int LNT = 256, LNW = LNT / 32, MAXP = 16;

struct LinPath { int off1, mul1, d, offo, mulo, woff, io, i1; float f; };
struct LinArgs {
    int npath, nio, nin, D1, Dout, nW;
    LinPath p[MAXP];
    long long rows;
    const float* x;      // forward: in1; gin: cotangent
    const float* y;
    const float* w;
    float* out;          // forward: out; gin: gin1
};

// forward: out[n][offo + w d + i]
__global__ void __launch_bounds__(LNT) o3lin_fwd_kernel(const __grid_constant__ LinArgs A) {
    extern __shared__ __align__(16) float lin_sm[];
    float* W = lin_sm;                                 // [path][u][w], scaled by f
    float* rows_s = lin_sm + A.nW;                     // [warp][D1]
    for (int k = 0; k < A.npath; ++k)
        for (int t = threadIdx.x; t < A.p[k].mul1 * A.p[k].mulo; t += LNT) W[A.p[k].woff + t] = A.p[k].f * __ldg(A.w + A.p[k].woff + t);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* xs = rows_s + warp * A.D1;
    for (long long n = (long long)blockIdx.x * LNW + warp; n < A.rows; n += (long long)gridDim.x * LNW) {
        __syncwarp();
        for (int c = lane; c < A.D1; c += 32) xs[c] = __ldg(A.x + n * A.D1 + c);
        const float y0 = __ldg(A.y + n);
        __syncwarp();
        float* o = A.out + n * A.Dout;
        for (int io = 0; io < A.nio; ++io) {
            int offo = -1, mulo = 0, d = 1;
            for (int k = 0; k < A.npath; ++k)
                if (A.p[k].io == io) { offo = A.p[k].offo; mulo = A.p[k].mulo; d = A.p[k].d; }
            if (offo < 0) continue;
            for (int e = lane; e < mulo * d; e += 32) {
                const int wch = e / d, i = e - wch * d;
                float acc = 0.f;
                for (int k = 0; k < A.npath; ++k) {
                    if (A.p[k].io != io) continue;
                    const float* wk = W + A.p[k].woff + wch;
                    const float* xk = xs + A.p[k].off1 + i;
                    for (int u = 0; u < A.p[k].mul1; ++u) acc = fmaf(wk[u * mulo], xk[u * d], acc);
                }
                o[offo + e] = y0 * acc;
            }
        }
    }
}

// input gradient: gx[n][off1 + u d + i] = y[n] * sum over the paths from i1 of f sum_w W[u][w] g[n][offo + w d + i]
__global__ void __launch_bounds__(LNT) o3lin_gin_kernel(const __grid_constant__ LinArgs A) {
    extern __shared__ __align__(16) float lin_sm[];
    float* Wt = lin_sm;                                // [path][w][u], scaled by f
    float* rows_s = lin_sm + A.nW;                     // [warp][Dout]
    for (int k = 0; k < A.npath; ++k) {
        const int m1 = A.p[k].mul1, mo = A.p[k].mulo;
        for (int t = threadIdx.x; t < m1 * mo; t += LNT) {
            const int u = t / mo, wch = t - u * mo;
            Wt[A.p[k].woff + wch * m1 + u] = A.p[k].f * __ldg(A.w + A.p[k].woff + t);
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* gs = rows_s + warp * A.Dout;
    for (long long n = (long long)blockIdx.x * LNW + warp; n < A.rows; n += (long long)gridDim.x * LNW) {
        __syncwarp();
        for (int c = lane; c < A.Dout; c += 32) gs[c] = __ldg(A.x + n * A.Dout + c);
        const float y0 = __ldg(A.y + n);
        __syncwarp();
        float* o = A.out + n * A.D1;
        for (int i1 = 0; i1 < A.nin; ++i1) {
            int off1 = -1, mul1 = 0, d = 1;
            for (int k = 0; k < A.npath; ++k)
                if (A.p[k].i1 == i1) { off1 = A.p[k].off1; mul1 = A.p[k].mul1; d = A.p[k].d; }
            if (off1 < 0) continue;      // an in1 irrep without a path: its gradient is zero (written below)
            for (int e = lane; e < mul1 * d; e += 32) {
                const int u = e / d, i = e - u * d;
                float acc = 0.f;
                for (int k = 0; k < A.npath; ++k) {
                    if (A.p[k].i1 != i1) continue;
                    const float* wk = Wt + A.p[k].woff + u;
                    const float* gk = gs + A.p[k].offo + i;
                    for (int wch = 0; wch < A.p[k].mulo; ++wch) acc = fmaf(wk[wch * mul1], gk[wch * d], acc);
                }
                o[off1 + e] = y0 * acc;
            }
        }
    }
}

}  // namespace

struct O3Lin {
    LinArgs A;
    std::vector<std::pair<int, int>> dead;   // (offset, width) of in1 irreps without a path
    size_t smem_f = 0, smem_g = 0;
};

O3Lin* o3lin_create(const o3::Plan& P) {
    if (getenv("SE3_O3TP_NO_LIN")) return nullptr;
    if (P.in2.size() != 1 || P.in2[0].l != 0 || (int)P.paths.size() > MAXP || P.nW > 11000 || P.D1 > 512 || P.Dout > 512) return nullptr;
    O3Lin* S = new O3Lin();
    LinArgs& A = S->A;
    memset(&A, 0, sizeof(A));
    std::vector<int> off1, offo;
    int acc = 0;
    for (auto& ir : P.in1) { off1.push_back(acc); acc += ir.mul * (2 * ir.l + 1); }
    acc = 0;
    for (auto& ir : P.out) { offo.push_back(acc); acc += ir.mul * (2 * ir.l + 1); }
    A.npath = (int)P.paths.size(); A.nio = (int)P.out.size(); A.nin = (int)P.in1.size(); A.D1 = P.D1; A.Dout = P.Dout; A.nW = P.nW;
    std::vector<char> has(P.in1.size(), 0);
    for (int k = 0; k < A.npath; ++k) {
        const o3::PathH& h = P.paths[k];
        const o3::Irrep a = P.in1[h.i1], o = P.out[h.io];
        if (a.l != o.l) { delete S; return nullptr; }
        const int d = 2 * a.l + 1;
        A.p[k] = {off1[h.i1], a.mul, d, offo[h.io], o.mul, h.woff, h.io, h.i1, P.a[h.io] / std::sqrt((float)d)};
        has[h.i1] = 1;
    }
    for (size_t i = 0; i < P.in1.size(); ++i)
        if (!has[i]) S->dead.push_back({off1[i], P.in1[i].mul * (2 * P.in1[i].l + 1)});
    // an output irrep without a path stays unwritten by the forward kernel: leave such plans to the general kernels
    std::vector<char> hit(P.out.size(), 0);
    for (auto& h : P.paths) hit[h.io] = 1;
    for (char c : hit)
        if (!c) { delete S; return nullptr; }
    if (!S->dead.empty()) { delete S; return nullptr; }
    S->smem_f = 4 * ((size_t)A.nW + (size_t)LNW * A.D1);
    S->smem_g = 4 * ((size_t)A.nW + (size_t)LNW * A.Dout);
    if (S->smem_f > 96 * 1024 || S->smem_g > 96 * 1024) { delete S; return nullptr; }
    return S;
}

void o3lin_destroy(O3Lin* S) { delete S; }

static int lin_launch(O3Lin* S, bool gin, long long rows, const float* x, const float* y, const float* w, float* out, cudaStream_t st) {
    LinArgs A = S->A;
    A.rows = rows; A.x = x; A.y = y; A.w = w; A.out = out;
    const size_t smem = gin ? S->smem_g : S->smem_f;
    static size_t attr_f = 48 * 1024, attr_g = 48 * 1024;
    size_t& attr = gin ? attr_g : attr_f;
    if (smem > attr) {
        if (gin) SE3_CUDA_TRY(cudaFuncSetAttribute(o3lin_gin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else SE3_CUDA_TRY(cudaFuncSetAttribute(o3lin_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    const int grid = (int)std::min<long long>((rows + LNW - 1) / LNW, (long long)se3::num_sms() * 8);
    if (gin) o3lin_gin_kernel<<<grid, LNT, smem, st>>>(A);
    else o3lin_fwd_kernel<<<grid, LNT, smem, st>>>(A);
    SE3_LAUNCHED();
    return SE3_OK;
}

int o3lin_forward(O3Lin* S, long long rows, const float* x, const float* y, const float* w, float* out, cudaStream_t st) {
    return lin_launch(S, false, rows, x, y, w, out, st);
}
int o3lin_gin(O3Lin* S, long long rows, const float* g, const float* y, const float* w, float* gx, cudaStream_t st) {
    return lin_launch(S, true, rows, g, y, w, gx, st);
}
