// First tensor product of the SEGNN message (msg1) by LINEARITY: node tables + per-edge SH combine.
//
// The reference op chain this replaces is L1TensorProduct.forward (L1TP:242-297) applied to
// in1 = cat(x[dst], x[src], edge_extra) with in2 = SH(1) of the edge.  For fixed in2 the product is linear in in1:
//     TP(cat(x_i, x_j, e), Y) = TP_dst(x_i, Y) + TP_src(x_j, Y) + TP_extra(e, Y),
// and every term factorises as "weight contraction of the raw channels" followed by "combination with Y":
//     P   = S  . [WZ_s | WV_s]            (scalars S of the node,   CH = mz + mv columns)
//     U_c = V_c. [WZ_d | WV_v]  c = x,y,z (vector components V_c)
//     out0[m]    = nz (Y0 P[m]      + c3 sum_c Y1[c] U_c[m])         (L1TP:242-256)
//     out1[m][c] = nv c3 (Y1[c] P[mz+m] + Y0 U_c[mz+m])              (L1TP:286-297)
// The contraction therefore runs ONCE PER NODE (a dense [Nn, D] x [D, 8 CH] GEMM over irrep channels, 16x fewer rows
// than edges at ~17 edges per node) into a table T[node][dst|src][CH][4] = (P, U_x, U_y, U_z) with norms and c3 folded
// in, and the per-edge work is a gather of two table rows, 4 multiply-adds per channel, and the gate.  Backward is
// the transpose: per-edge gate VJP -> Y^T combine -> segment sums into G[node][dst|src][CH][4], then node-level GEMMs.
//
// Kernels here (all SIMT fp32, HBM/L2-bound streaming work; the weight contraction itself is the node-level GEMM):
//   msg1_edge_fwd      one warp per destination node (its CSR row): the dst table row stays in registers, two edges
//                      per step (half-warps), lanes = channels -> coalesced 16-byte table loads; writes the
//                      pre-activation [E, ns+4nv] and the gated message [E, ns+3nv]
//   msg1_edge_bwd_dst  same mapping: gate VJP -> g_pre [E, ns+4nv] (kept for the src pass), segment sum over the row
//                      in registers -> G dst half (no atomics, no zero-init), extras' weight gradient per warp
//   msg1_edge_bwd_src  one warp per SOURCE node over the transposed edge order (graph_transpose below): G src half
//   msg1_expand / msg1_contract  weights <-> the dense [D, 8 CH] matrix of the node GEMM (and its gradient)
//   graph_transpose / rowptr_from_sorted  stable counting sort of the edges by source, CSR row pointers of a sorted index
#include <algorithm>

#include "tc_common.cuh"

namespace se3 {

template <int NS, int NV>
struct MsgDims {
    static constexpr int MZ = NS + NV;          // 0e outputs of the product: scalars + one gate per vector channel
    static constexpr int CH = NS + 2 * NV;      // output channels: MZ 0e + NV 1o
    static constexpr int DPRE = NS + 4 * NV;    // pre-activation row  [MZ | NV x 3]
    static constexpr int DPOST = NS + 3 * NV;   // gated row           [NS | NV x 3]
    static constexpr int D = NS + 3 * NV;       // node feature row
    static constexpr int HALF = 4 * CH;         // floats of one table half (dst or src)
    static constexpr int LDT = 2 * HALF;
    static constexpr int NG = (MZ + 15) / 16;   // 16-lane channel groups
};

struct Msg1Fwd {
    long long n_dst;
    const long long* rowptr;
    const int* src;
    const float* table;      // [n_all, LDT]
    const float* we;         // [2, CH] extras' weights (P column)
    const float* y;          // [E, 4]
    const float* extra;      // [E, 2]
    float* pre;              // [E, DPRE]
    float* post;             // [E, DPOST]
    float cs, cg;
};

struct Msg1Bwd {
    long long n_dst, n_all;
    const long long* rowptr;
    const long long* tptr;
    const int* perm;
    const float* y;
    const float* extra;
    const float* pre;
    const float* gpost;
    float* gpre;
    float* G;                // [n_all, LDT]
    float* gwe_part;         // [gridDim.x, 2, CH]
    float cs, cg;
};

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float2 ld2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }

template <int NS, int NV>
__global__ void __launch_bounds__(256) msg1_edge_fwd_kernel(const Msg1Fwd A) {
    using Dm = MsgDims<NS, NV>;
    constexpr int MZ = Dm::MZ, CH = Dm::CH, NG = Dm::NG;
    const int lane = threadIdx.x & 31, half = lane >> 4, l16 = lane & 15;
    const long long w0 = (long long)blockIdx.x * 8 + (threadIdx.x >> 5), wstride = (long long)gridDim.x * 8;
    float we0[NG], we1[NG], wv0[NG], wv1[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const int ch = g * 16 + l16;
        we0[g] = ch < MZ ? __ldg(A.we + ch) : 0.0f;
        we1[g] = ch < MZ ? __ldg(A.we + CH + ch) : 0.0f;
        const bool isv = ch >= NS && ch < MZ;
        wv0[g] = isv ? __ldg(A.we + MZ + ch - NS) : 0.0f;
        wv1[g] = isv ? __ldg(A.we + CH + MZ + ch - NS) : 0.0f;
    }
    for (long long n = w0; n < A.n_dst; n += wstride) {
        const long long e0 = __ldg(A.rowptr + n), e1 = __ldg(A.rowptr + n + 1);
        if (e0 == e1) continue;
        const float* tn = A.table + n * Dm::LDT;
        float4 td[NG], tv[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const int ch = g * 16 + l16;
            td[g] = ch < MZ ? ld4(tn + 4 * ch) : make_float4(0.f, 0.f, 0.f, 0.f);
            tv[g] = (ch >= NS && ch < MZ) ? ld4(tn + 4 * (MZ + ch - NS)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (long long e = e0 + half; e < e1; e += 2) {
            const int s = __ldg(A.src + e);
            const float4 y = ld4(A.y + 4 * e);
            const float2 ex = ld2(A.extra + 2 * e);
            const float* ts = A.table + (long long)s * Dm::LDT + Dm::HALF;
            float* pre = A.pre + e * Dm::DPRE;
            float* post = A.post + e * Dm::DPOST;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const int ch = g * 16 + l16;
                if (ch < MZ) {
                    const float4 t = ld4(ts + 4 * ch);
                    const float P = td[g].x + t.x + fmaf(ex.x, we0[g], ex.y * we1[g]);
                    const float x = fmaf(y.x, P, fmaf(y.y, td[g].y + t.y, fmaf(y.z, td[g].z + t.z, y.w * (td[g].w + t.w))));
                    pre[ch] = x;
                    const float sg = sigm(x);
                    if (ch < NS) {
                        post[ch] = A.cs * x * sg;
                    } else {
                        const int v = ch - NS;
                        const float4 t2 = ld4(ts + 4 * (MZ + v));
                        const float Pv = tv[g].x + t2.x + fmaf(ex.x, wv0[g], ex.y * wv1[g]);
                        const float p0 = fmaf(y.y, Pv, y.x * (tv[g].y + t2.y));
                        const float p1 = fmaf(y.z, Pv, y.x * (tv[g].z + t2.z));
                        const float p2 = fmaf(y.w, Pv, y.x * (tv[g].w + t2.w));
                        const float gs = A.cg * sg;
                        float* pv = pre + MZ + 3 * v;
                        pv[0] = p0; pv[1] = p1; pv[2] = p2;
                        float* qv = post + NS + 3 * v;
                        qv[0] = gs * p0; qv[1] = gs * p1; qv[2] = gs * p2;
                    }
                }
            }
        }
    }
}

// Backward segment sums, one warp per node, ONE EDGE PER WARP STEP with every lane busy: lanes [0, MZ/2) own two 0e
// channels each (scalars or gates), lanes [MZ/2, MZ/2 + NV) one vector channel each (34x0e+10x1o: 22 + 10 = 32 lanes).
// A lane keeps the (P, U_x, U_y, U_z) sums of its channels in registers for the whole CSR row and writes them once.
//   MODE 0: rows of the CSR by destination; gate VJP from (pre, gpost) -> gpre (kept for the src pass), dst half of G,
//           extras' weight gradient
//   MODE 1: segments of the transposed (by source) order; reads gpre through perm, src half of G
//   MODE 2: as MODE 0 but gpre is an input (the gate VJP was done by the producer of gpre)
template <int NS, int NV, int MODE>
__global__ void __launch_bounds__(256) msg1_edge_bwd_kernel(const Msg1Bwd A) {
    using Dm = MsgDims<NS, NV>;
    constexpr int MZ = Dm::MZ, CH = Dm::CH, ZL = MZ / 2;
    static_assert(MZ % 2 == 0 && NS % 2 == 0 && ZL + NV <= 32, "lane mapping: two 0e channels or one vector channel per lane");
    constexpr bool SRC = MODE == 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long w0 = (long long)blockIdx.x * 8 + warp, wstride = (long long)gridDim.x * 8;
    const bool zl = lane < ZL, vl = lane >= ZL && lane < ZL + NV;
    const bool gl = zl && 2 * lane >= NS;                 // gate lane: gates v0 = 2 lane - NS, v0 + 1
    const int v = lane - ZL;                              // vector lanes: channel v
    const int c0 = 2 * lane;                              // z lanes: channels c0, c0 + 1
    const int gsrc = vl ? (NS + v) >> 1 : 0, gsel = vl ? (NS + v) & 1 : 0;     // lane / element holding the gate of v
    const int d0 = gl ? ZL + (c0 - NS) : 0;               // gate lanes: vector lanes of their two gates (d0, d0 + 1)
    float ge[4] = {0.f, 0.f, 0.f, 0.f};                   // extras' gradient: z lanes [j][c0 + i] -> ge[2 j + i]; vector lanes ge[j]
    const long long* ptr = SRC ? A.tptr : A.rowptr;
    const long long nseg = SRC ? A.n_all : A.n_dst;
    for (long long n = w0; n < A.n_all; n += wstride) {
        long long e0 = 0, e1 = 0;
        if (n < nseg) { e0 = __ldg(ptr + n); e1 = __ldg(ptr + n + 1); }
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;       // z lanes: channels c0, c0 + 1; vector lanes: a0 only
        if (MODE != 0) {
            // gpre is an input: four edges per step, all loads issued before the first use (the loop is latency-bound)
            for (long long k = e0; k < e1; k += 4) {
                long long ee[4];
                float4 yy[4];
                float2 xx[4], gz[4];
                float qa[4], qb[4], qc[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const long long kk = k + j < e1 ? k + j : e1 - 1;
                    ee[j] = SRC ? (long long)__ldg(A.perm + kk) : kk;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    yy[j] = ld4(A.y + 4 * ee[j]);
                    xx[j] = SRC ? make_float2(0.f, 0.f) : ld2(A.extra + 2 * ee[j]);
                    const float* gp = A.gpre + ee[j] * Dm::DPRE;
                    gz[j] = zl ? ld2(gp + c0) : make_float2(0.f, 0.f);
                    const float* qv = gp + MZ + 3 * (vl ? v : 0);
                    qa[j] = vl ? __ldg(qv) : 0.f; qb[j] = vl ? __ldg(qv + 1) : 0.f; qc[j] = vl ? __ldg(qv + 2) : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (k + j >= e1) break;
                    const float4 y = yy[j];
                    if (zl) {
                        const float p0 = y.x * gz[j].x, p1 = y.x * gz[j].y;
                        a0.x += p0; a0.y = fmaf(y.y, gz[j].x, a0.y); a0.z = fmaf(y.z, gz[j].x, a0.z); a0.w = fmaf(y.w, gz[j].x, a0.w);
                        a1.x += p1; a1.y = fmaf(y.y, gz[j].y, a1.y); a1.z = fmaf(y.z, gz[j].y, a1.z); a1.w = fmaf(y.w, gz[j].y, a1.w);
                        if (!SRC) {
                            ge[0] = fmaf(xx[j].x, p0, ge[0]); ge[1] = fmaf(xx[j].x, p1, ge[1]);
                            ge[2] = fmaf(xx[j].y, p0, ge[2]); ge[3] = fmaf(xx[j].y, p1, ge[3]);
                        }
                    } else if (vl) {
                        const float d = fmaf(y.y, qa[j], fmaf(y.z, qb[j], y.w * qc[j]));
                        a0.x += d; a0.y = fmaf(y.x, qa[j], a0.y); a0.z = fmaf(y.x, qb[j], a0.z); a0.w = fmaf(y.x, qc[j], a0.w);
                        if (!SRC) { ge[0] = fmaf(xx[j].x, d, ge[0]); ge[1] = fmaf(xx[j].y, d, ge[1]); }
                    }
                }
            }
        } else
        for (long long k = e0; k < e1; ++k) {
            const long long e = SRC ? (long long)__ldg(A.perm + k) : k;
            const float4 y = ld4(A.y + 4 * e);
            float2 ex = make_float2(0.f, 0.f);
            if (!SRC) ex = ld2(A.extra + 2 * e);
            float* gp = A.gpre + e * Dm::DPRE;
            float gx0 = 0.f, gx1 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f;
            if (MODE == 0) {
                const float* pre = A.pre + e * Dm::DPRE;
                const float* gm = A.gpost + e * Dm::D;
                float x0 = 0.f, x1 = 0.f, sg0 = 0.f, sg1 = 0.f, dot = 0.f, g0 = 0.f, g1 = 0.f, g2 = 0.f;
                if (zl) {
                    const float2 x = ld2(pre + c0);
                    x0 = x.x; x1 = x.y;
                    sg0 = sigm(x0); sg1 = sigm(x1);
                }
                if (vl) {
                    const float* gv = gm + NS + 3 * v;
                    const float* pv = pre + MZ + 3 * v;
                    g0 = __ldg(gv); g1 = __ldg(gv + 1); g2 = __ldg(gv + 2);
                    dot = fmaf(g0, __ldg(pv), fmaf(g1, __ldg(pv + 1), g2 * __ldg(pv + 2)));
                }
                // gate sigmoid -> vector lanes; vector dots -> gate lanes
                const float s_a = __shfl_sync(0xffffffffu, sg0, gsrc), s_b = __shfl_sync(0xffffffffu, sg1, gsrc);
                const float dt0 = __shfl_sync(0xffffffffu, dot, d0), dt1 = __shfl_sync(0xffffffffu, dot, d0 + 1);
                if (zl) {
                    if (!gl) {
                        const float2 g = ld2(gm + c0);
                        gx0 = A.cs * g.x * sg0 * fmaf(x0, 1.0f - sg0, 1.0f);
                        gx1 = A.cs * g.y * sg1 * fmaf(x1, 1.0f - sg1, 1.0f);
                    } else {
                        gx0 = A.cg * sg0 * (1.0f - sg0) * dt0;
                        gx1 = A.cg * sg1 * (1.0f - sg1) * dt1;
                    }
                    *reinterpret_cast<float2*>(gp + c0) = make_float2(gx0, gx1);
                }
                if (vl) {
                    const float gs = A.cg * (gsel ? s_b : s_a);
                    q0 = gs * g0; q1 = gs * g1; q2 = gs * g2;
                    float* qo = gp + MZ + 3 * v;
                    qo[0] = q0; qo[1] = q1; qo[2] = q2;
                }
            } else {
                if (zl) { const float2 g = ld2(gp + c0); gx0 = g.x; gx1 = g.y; }
                if (vl) { const float* qv = gp + MZ + 3 * v; q0 = __ldg(qv); q1 = __ldg(qv + 1); q2 = __ldg(qv + 2); }
            }
            if (zl) {
                const float p0 = y.x * gx0, p1 = y.x * gx1;
                a0.x += p0; a0.y = fmaf(y.y, gx0, a0.y); a0.z = fmaf(y.z, gx0, a0.z); a0.w = fmaf(y.w, gx0, a0.w);
                a1.x += p1; a1.y = fmaf(y.y, gx1, a1.y); a1.z = fmaf(y.z, gx1, a1.z); a1.w = fmaf(y.w, gx1, a1.w);
                if (!SRC) {
                    ge[0] = fmaf(ex.x, p0, ge[0]); ge[1] = fmaf(ex.x, p1, ge[1]);
                    ge[2] = fmaf(ex.y, p0, ge[2]); ge[3] = fmaf(ex.y, p1, ge[3]);
                }
            } else if (vl) {
                const float d = fmaf(y.y, q0, fmaf(y.z, q1, y.w * q2));
                a0.x += d; a0.y = fmaf(y.x, q0, a0.y); a0.z = fmaf(y.x, q1, a0.z); a0.w = fmaf(y.x, q2, a0.w);
                if (!SRC) { ge[0] = fmaf(ex.x, d, ge[0]); ge[1] = fmaf(ex.y, d, ge[1]); }
            }
        }
        float* gn = A.G + n * Dm::LDT + (SRC ? Dm::HALF : 0);
        if (zl) {
            *reinterpret_cast<float4*>(gn + 4 * c0) = a0;
            *reinterpret_cast<float4*>(gn + 4 * c0 + 4) = a1;
        } else if (vl) {
            *reinterpret_cast<float4*>(gn + 4 * (MZ + v)) = a0;
        }
    }
    if (!SRC) {
        // extras' weight gradient of this block: [2][CH] partial (summed deterministically by msg1_contract)
        __shared__ float sgw[8][2][CH];
        if (zl) {
            sgw[warp][0][c0] = ge[0]; sgw[warp][0][c0 + 1] = ge[1];
            sgw[warp][1][c0] = ge[2]; sgw[warp][1][c0 + 1] = ge[3];
        } else if (vl) {
            sgw[warp][0][MZ + v] = ge[0]; sgw[warp][1][MZ + v] = ge[1];
        }
        __syncthreads();
        for (int t = threadIdx.x; t < 2 * CH; t += 256) {
            float s = 0.0f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += (&sgw[w][0][0])[t];
            A.gwe_part[(long long)blockIdx.x * 2 * CH + t] = s;
        }
    }
}

// wbig [D, LDT]: row = node feature column (NS scalars, then NV x (x,y,z)); column = half * HALF + 4 ch + j.
// we [2, CH].  Weight rows follow the concatenation order of the reference (L1TP:81-88): all 0e channels
// (dst scalars, src scalars, extras), then all 1o channels (dst vectors, src vectors).
template <int NS, int NV>
__global__ void msg1_expand_kernel(const float* __restrict__ wz, const float* __restrict__ wv, const float* __restrict__ nz,
                                   const float* __restrict__ nvn, float* __restrict__ wbig, float* __restrict__ we) {
    using Dm = MsgDims<NS, NV>;
    constexpr int MZ = Dm::MZ, CH = Dm::CH, NSC = 2 * NS + 2;
    const int total = Dm::D * Dm::LDT;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total + 2 * CH; t += gridDim.x * blockDim.x) {
        if (t >= total) {
            const int u = t - total, j = u / CH, ch = u - j * CH;
            const int row = 2 * NS + j;
            we[u] = ch < MZ ? (nz ? nz[ch] : 1.0f) * wz[row * MZ + ch]
                            : (nvn ? nvn[3 * (ch - MZ)] : 1.0f) * C3f * wv[row * NV + ch - MZ];
            continue;
        }
        const int k = t / Dm::LDT, c = t - k * Dm::LDT;
        const int p = c / Dm::HALF, cc = c - p * Dm::HALF, ch = cc >> 2, j = cc & 3;
        int row = -1;
        float f = 1.0f;
        if (k < NS) { if (j == 0) row = p * NS + k; }
        else {
            const int kv = (k - NS) / 3, comp = (k - NS) - 3 * kv;
            if (j == 1 + comp) { row = NSC + p * NV + kv; f = C3f; }
        }
        float val = 0.0f;
        if (row >= 0) {
            if (ch < MZ) val = f * (nz ? nz[ch] : 1.0f) * wz[row * MZ + ch];
            else val = C3f * (nvn ? nvn[3 * (ch - MZ)] : 1.0f) * wv[row * NV + ch - MZ];
        }
        if (wbig) wbig[t] = val;
    }
}

// transpose of msg1_expand: gwz [(2 NS + 2 + 2 NV), MZ], gwv [(same), NV] overwritten from gwbig [D, LDT] and the
// per-block partials of the extras' gradient gwe_part [nparts, 2, CH].
template <int NS, int NV>
__global__ void msg1_contract_kernel(const float* __restrict__ gwbig, const float* __restrict__ gwe_part, int nparts,
                                     const float* __restrict__ nz, const float* __restrict__ nvn, float* __restrict__ gwz,
                                     float* __restrict__ gwv) {
    using Dm = MsgDims<NS, NV>;
    constexpr int MZ = Dm::MZ, CH = Dm::CH, NSC = 2 * NS + 2, ROWS = NSC + 2 * NV;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ROWS * CH; t += gridDim.x * blockDim.x) {
        const int row = t / CH, ch = t - row * CH;
        float g = 0.0f;
        if (row < 2 * NS) {
            const int p = row / NS, k = row - p * NS;
            g = gwbig[k * Dm::LDT + p * Dm::HALF + 4 * ch];
        } else if (row < NSC) {
            const int j = row - 2 * NS;
            for (int q = 0; q < nparts; ++q) g += gwe_part[((long long)q * 2 + j) * CH + ch];
        } else {
            const int r = row - NSC, p = r / NV, kv = r - p * NV;
#pragma unroll
            for (int c = 0; c < 3; ++c) g += gwbig[(NS + 3 * kv + c) * Dm::LDT + p * Dm::HALF + 4 * ch + 1 + c];
            if (ch < MZ) g *= C3f;
        }
        if (ch < MZ) gwz[row * MZ + ch] = g * (nz ? nz[ch] : 1.0f);
        else gwv[row * NV + ch - MZ] = g * C3f * (nvn ? nvn[3 * (ch - MZ)] : 1.0f);
    }
}

// ------------------------------------------------------------------ graph helpers
__global__ void rowptr_fill_kernel(const int* __restrict__ idx, long long e, long long n, long long* __restrict__ rowptr) {
    // rowptr[k] = first position whose idx >= k  (idx sorted ascending); rowptr[n] = e
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > e) return;
    const long long lo = i == 0 ? 0 : (long long)idx[i - 1] + 1;      // nodes (idx[i-1], idx[i]] start at i
    const long long hi = i == e ? n : (long long)idx[i];
    for (long long k = lo; k <= hi && k <= n; ++k) rowptr[k] = i;
}

__global__ void tr_count_kernel(const int* __restrict__ src, long long e, int* __restrict__ cnt) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < e) atomicAdd(cnt + src[i], 1);
}
// single-block exclusive scan of int counts into int64 pointers (n + 1 entries); also seeds the fill cursors.
// Chunks of 1024 consecutive elements (coalesced), warp-shuffle scan per chunk, running carry.
__global__ void __launch_bounds__(1024) tr_scan_kernel(const int* __restrict__ cnt, long long n, long long* __restrict__ ptr,
                                                        unsigned long long* __restrict__ cursor) {
    __shared__ long long wsum[32];
    __shared__ long long carry_s;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (long long base = 0; base < n; base += 1024) {
        const long long i = base + t;
        const long long v = i < n ? (long long)cnt[i] : 0;
        long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        long long off = carry_s;
        for (int w = 0; w < warp; ++w) off += wsum[w];
        if (i < n) { const long long ex = off + inc - v; ptr[i] = ex; cursor[i] = (unsigned long long)ex; }
        __syncthreads();
        if (t == 1023) carry_s = off + inc;
        __syncthreads();
    }
    if (t == 0) ptr[n] = carry_s;
}
// Multi-block version of the scan (the single block above walks 1 100 chunks for the nodes of a 1M-particle cloud:
// 1.4 ms): per-1024-chunk sums, one block scans the chunk sums, every chunk finishes its own scan with its offset.
__device__ __forceinline__ long long tr_block_scan_1024(long long v, long long* wsum /*[32]*/, long long& total) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    long long off = 0, tot = 0;
    for (int w = 0; w < 32; ++w) {
        const long long x = wsum[w];
        if (w < warp) off += x;
        tot += x;
    }
    total = tot;
    __syncthreads();
    return off + inc - v;     // exclusive
}
__global__ void __launch_bounds__(1024) tr_bsum_kernel(const int* __restrict__ cnt, long long n, long long* __restrict__ bsum) {
    __shared__ long long wsum[32];
    const long long i = (long long)blockIdx.x * 1024 + threadIdx.x;
    long long tot;
    tr_block_scan_1024(i < n ? (long long)cnt[i] : 0, wsum, tot);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(1024) tr_btop_kernel(long long* __restrict__ bsum, long long nb) {
    __shared__ long long wsum[32];
    long long carry = 0;
    for (long long base = 0; base < nb; base += 1024) {
        const long long i = base + threadIdx.x;
        long long tot;
        const long long ex = tr_block_scan_1024(i < nb ? bsum[i] : 0, wsum, tot);
        if (i < nb) bsum[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) bsum[nb] = carry;
}
__global__ void __launch_bounds__(1024) tr_final_kernel(const int* __restrict__ cnt, long long n, const long long* __restrict__ bsum,
                                                         long long nb, long long* __restrict__ ptr,
                                                         unsigned long long* __restrict__ cursor) {
    __shared__ long long wsum[32];
    const long long i = (long long)blockIdx.x * 1024 + threadIdx.x;
    long long tot;
    const long long ex = bsum[blockIdx.x] + tr_block_scan_1024(i < n ? (long long)cnt[i] : 0, wsum, tot);
    if (i < n) { ptr[i] = ex; cursor[i] = (unsigned long long)ex; }
    if (blockIdx.x == 0 && threadIdx.x == 0) ptr[n] = bsum[nb];
}
__global__ void tr_fill_kernel(const int* __restrict__ src, long long e, unsigned long long* __restrict__ cursor, int* __restrict__ perm) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < e) perm[atomicAdd(cursor + src[i], 1ull)] = (int)i;
}
// the fill order inside a segment is arbitrary: sort every segment ascending by edge id (= stable by source: the
// summation order of the src pass, and with it the result, is run-to-run deterministic).  One warp per segment:
// segments of up to 64 edges (every node of the octree graph: <= 32 + 26 + 8 + 1 neighbours) by rank counting with
// shuffles (coalesced load, rank = number of smaller ids, scattered store), longer ones by a serial insertion sort.
__global__ void __launch_bounds__(256) tr_sort_kernel(const long long* __restrict__ ptr, long long n, int* __restrict__ perm) {
    const int lane = threadIdx.x & 31;
    const long long j = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j >= n) return;
    const long long a = ptr[j], b = ptr[j + 1];
    const int len = (int)(b - a);
    if (len <= 1) return;
    if (len <= 64) {
        const int v0 = lane < len ? perm[a + lane] : 0x7fffffff;
        const int v1 = 32 + lane < len ? perm[a + 32 + lane] : 0x7fffffff;
        int r0 = 0, r1 = 0;
        const int l0 = len < 32 ? len : 32;
        if (len <= 32) {
            for (int l = 0; l < l0; ++l) r0 += __shfl_sync(0xffffffffu, v0, l) < v0;
        } else {
#pragma unroll 8
            for (int l = 0; l < 32; ++l) {
                const int u0 = __shfl_sync(0xffffffffu, v0, l), u1 = __shfl_sync(0xffffffffu, v1, l);
                r0 += (u0 < v0) + (u1 < v0);
                r1 += (u0 < v1) + (u1 < v1);
            }
        }
        __syncwarp();
        if (lane < len) perm[a + r0] = v0;          // edge ids are distinct: ranks are a permutation
        if (32 + lane < len) perm[a + r1] = v1;
    } else if (lane == 0) {
        for (long long i = a + 1; i < b; ++i) {
            const int v = perm[i];
            long long k = i;
            while (k > a && perm[k - 1] > v) { perm[k] = perm[k - 1]; --k; }
            perm[k] = v;
        }
    }
}

template <int NS, int NV>
static int msg1_launch_fwd(const Msg1Fwd& A, cudaStream_t st) {
    const int grid = (int)std::max<long long>(1, std::min<long long>((A.n_dst + 7) / 8, (long long)num_sms() * 8));
    msg1_edge_fwd_kernel<NS, NV><<<grid, 256, 0, st>>>(A);
    SE3_LAUNCHED();
    return SE3_OK;
}
template <int NS, int NV>
static int msg1_launch_bwd(const Msg1Bwd& A, int grid, cudaStream_t st) {
    if (A.pre) msg1_edge_bwd_kernel<NS, NV, 0><<<grid, 256, 0, st>>>(A);
    else msg1_edge_bwd_kernel<NS, NV, 2><<<grid, 256, 0, st>>>(A);    // gpre given: the gate VJP was done by its producer
    SE3_LAUNCHED();
    msg1_edge_bwd_kernel<NS, NV, 1><<<grid, 256, 0, st>>>(A);
    SE3_LAUNCHED();
    return SE3_OK;
}

}  // namespace se3

using namespace se3;

#define SE3_MSG1_DISPATCH(ns, nv, CALL)                                      \
    if ((ns) == 34 && (nv) == 10) { CALL(34, 10) }                           \
    else if ((ns) == 16 && (nv) == 8) { CALL(16, 8) }                        \
    else if ((ns) == 8 && (nv) == 4) { CALL(8, 4) }                          \
    else { set_error("msg1: hidden irreps %dx0e+%dx1o are not instantiated", (int)(ns), (int)(nv)); return SE3_ERR_INVALID; }

extern "C" int se3_msg1_supported(int32_t ns, int32_t nv, int32_t ne) {
    return ne == 2 && ((ns == 34 && nv == 10) || (ns == 16 && nv == 8) || (ns == 8 && nv == 4)) ? 1 : 0;
}

extern "C" int se3_msg1_max_parts(void) { return num_sms() * 8; }

extern "C" int se3_msg1_expand(int32_t ns, int32_t nv, const float* wz, const float* wv, const float* nz, const float* nvn,
                               float* wbig, float* we, void* stream) {
    if (!wz || !wv || !we) { set_error("msg1_expand: null argument"); return SE3_ERR_INVALID; }   // wbig may be NULL
#define CALL(a, b) msg1_expand_kernel<a, b><<<64, 256, 0, (cudaStream_t)stream>>>(wz, wv, nz, nvn, wbig, we);
    SE3_MSG1_DISPATCH(ns, nv, CALL)
#undef CALL
    SE3_LAUNCHED();
    return SE3_OK;
}

extern "C" int se3_msg1_contract(int32_t ns, int32_t nv, const float* gwbig, const float* gwe_part, int32_t nparts,
                                 const float* nz, const float* nvn, float* gwz, float* gwv, void* stream) {
    if (!gwbig || !gwe_part || !gwz || !gwv || nparts < 0) { set_error("msg1_contract: bad argument"); return SE3_ERR_INVALID; }
#define CALL(a, b) msg1_contract_kernel<a, b><<<32, 256, 0, (cudaStream_t)stream>>>(gwbig, gwe_part, nparts, nz, nvn, gwz, gwv);
    SE3_MSG1_DISPATCH(ns, nv, CALL)
#undef CALL
    SE3_LAUNCHED();
    return SE3_OK;
}

extern "C" int se3_msg1_edge_forward(int32_t ns, int32_t nv, int64_t n_dst, const int64_t* rowptr, const int32_t* src,
                                     const float* table, const float* we, const float* y, const float* extra, float cs,
                                     float cg, float* pre, float* post, void* stream) {
    if (n_dst < 0) { set_error("msg1_edge_forward: bad argument"); return SE3_ERR_INVALID; }
    if (n_dst == 0) return SE3_OK;
    if (!rowptr || !src || !table || !we || !y || !extra || !pre || !post) { set_error("msg1_edge_forward: null argument"); return SE3_ERR_INVALID; }
    Msg1Fwd A;
    A.n_dst = n_dst; A.rowptr = (const long long*)rowptr; A.src = src; A.table = table; A.we = we; A.y = y; A.extra = extra;
    A.pre = pre; A.post = post; A.cs = cs; A.cg = cg;
#define CALL(a, b) return msg1_launch_fwd<a, b>(A, (cudaStream_t)stream);
    SE3_MSG1_DISPATCH(ns, nv, CALL)
#undef CALL
}

extern "C" int se3_msg1_edge_backward(int32_t ns, int32_t nv, int64_t n_dst, int64_t n_all, const int64_t* rowptr,
                                      const int64_t* tptr, const int32_t* perm, const float* y, const float* extra,
                                      const float* pre, const float* gpost, float cs, float cg, float* gpre, float* G,
                                      float* gwe_part, int32_t* nparts, void* stream) {
    if (n_dst < 0 || n_all < n_dst || !nparts) { set_error("msg1_edge_backward: bad argument"); return SE3_ERR_INVALID; }
    *nparts = 0;
    if (n_all == 0) return SE3_OK;
    if (!rowptr || !tptr || !perm || !y || !extra || !gpre || !G || !gwe_part || (pre && !gpost)) {
        set_error("msg1_edge_backward: null argument");
        return SE3_ERR_INVALID;
    }
    Msg1Bwd A;
    A.n_dst = n_dst; A.n_all = n_all; A.rowptr = (const long long*)rowptr; A.tptr = (const long long*)tptr; A.perm = perm;
    A.y = y; A.extra = extra; A.pre = pre; A.gpost = gpost; A.gpre = gpre; A.G = G; A.gwe_part = gwe_part; A.cs = cs; A.cg = cg;
    const int grid = (int)std::max<long long>(1, std::min<long long>((n_all + 7) / 8, (long long)num_sms() * 8));
    *nparts = grid;
#define CALL(a, b) return msg1_launch_bwd<a, b>(A, grid, (cudaStream_t)stream);
    SE3_MSG1_DISPATCH(ns, nv, CALL)
#undef CALL
}

extern "C" int se3_rowptr_from_sorted(int64_t e, int64_t n, const int32_t* idx_sorted, int64_t* rowptr, void* stream) {
    if (e < 0 || n < 0 || !rowptr || (e > 0 && !idx_sorted)) { set_error("rowptr_from_sorted: bad argument"); return SE3_ERR_INVALID; }
    const unsigned grid = (unsigned)((e + 1 + 255) / 256);
    rowptr_fill_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(idx_sorted, e, n, (long long*)rowptr);
    SE3_LAUNCHED();
    return SE3_OK;
}

extern "C" int se3_graph_transpose_work_bytes(int64_t n_src, size_t* bytes) {
    if (n_src < 0 || !bytes) { set_error("graph_transpose: bad argument"); return SE3_ERR_INVALID; }
    *bytes = (size_t)(n_src + 1) * (sizeof(int) + sizeof(unsigned long long)) + ((size_t)(n_src + 1023) / 1024 + 2) * sizeof(long long) + 256;
    return SE3_OK;
}

extern "C" int se3_graph_transpose(int64_t e, int64_t n_src, const int32_t* src, int64_t* tptr, int32_t* perm, void* work,
                                   size_t work_bytes, void* stream) {
    size_t need = 0;
    if (se3_graph_transpose_work_bytes(n_src, &need) || e < 0 || !tptr || !work || work_bytes < need || (e > 0 && (!src || !perm))) {
        set_error("graph_transpose: bad argument / workspace too small");
        return SE3_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* cursor = (unsigned long long*)work;
    int* cnt = (int*)(cursor + n_src + 1);
    SE3_CUDA_TRY(cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)(n_src + 1), st));
    if (e > 0) { tr_count_kernel<<<(unsigned)((e + 255) / 256), 256, 0, st>>>(src, e, cnt); SE3_LAUNCHED(); }
    const long long nb = (n_src + 1023) / 1024;
    if (nb <= 8) {
        tr_scan_kernel<<<1, 1024, 0, st>>>(cnt, n_src, (long long*)tptr, cursor); SE3_LAUNCHED();
    } else {
        // 8-byte aligned chunk sums behind the counts
        long long* bsum = (long long*)(((uintptr_t)(cnt + n_src + 1) + 7) & ~(uintptr_t)7);
        tr_bsum_kernel<<<(unsigned)nb, 1024, 0, st>>>(cnt, n_src, bsum); SE3_LAUNCHED();
        tr_btop_kernel<<<1, 1024, 0, st>>>(bsum, nb); SE3_LAUNCHED();
        tr_final_kernel<<<(unsigned)nb, 1024, 0, st>>>(cnt, n_src, bsum, nb, (long long*)tptr, cursor); SE3_LAUNCHED();
    }
    if (e > 0) {
        tr_fill_kernel<<<(unsigned)((e + 255) / 256), 256, 0, st>>>(src, e, cursor, perm); SE3_LAUNCHED();
        tr_sort_kernel<<<(unsigned)((n_src + 7) / 8), 256, 0, st>>>((const long long*)tptr, n_src, perm); SE3_LAUNCHED();
    }
    return SE3_OK;
}
