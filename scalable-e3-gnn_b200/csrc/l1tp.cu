// l<=1 Clebsch-Gordan tensor product with SH(1) second input, fully-connected weights.
//
// Replaces L1TensorProduct.forward (/root/reference/models/segnn/l1_tensor_prod.py:234-299)
// and its autograd backward with two persistent, tile-based sm_100a kernels.
//
// Reference formulation (L1TP:244-297), per row r with Y0=in2[r,0], Y1=in2[r,1:4]:
//     out0e = ([s0e*Y0 , c3<v1o,Y1>] @ W0e) * norm      out1o = ([c3 s0e(x)Y1, c3 v1o*Y0, c6 v1e x Y1] @ W1o) * norm
//     out0o = ([s0o*Y0 , c3<v1e,Y1>] @ W0o) * norm      out1e = ([c3 s0o(x)Y1, c3 v1e*Y0, c6 v1o x Y1] @ W1e) * norm
// Here the two parity "families" (E: s0e,v1o,v1e -> 0e,1o ; O: s0o,v1e,v1o -> 0o,1e) share code, and the
// per-row scalars Y0 / Y1[c] that multiply *scalar* inputs are factored out of the K-loop:
//     outZ[m]    = norm * ( Y0 * sum_k s_k WZ[k,m]  +  sum_k (c3<v_k,Y1>) WZ[ns+k,m] )
//     outV[m][c] = norm * ( c3 Y1[c] * sum_k s_k WV[k,m]  +  sum_k A_V[(r,c)][k] WV[ns+k,m] )
// which removes 2/3 of the l=1 scalar-path FMAs and the [E,3,K] feature tensor of the reference.
//
// Tile pipeline (one CTA = 256 threads, TR rows per tile, grid-stride over tiles):
//   gather rows (virtual concat of up to 4 indexed segments) -> feature tiles AZ/AV in shared memory
//   -> register-tiled fp32 GEMMs against weights resident in shared memory (4 rows x TN cols per lane,
//      lanes = 8 row groups x 4 column groups, all operand loads are conflict-free LDS.128)
//   -> out tile in shared memory -> fused epilogue (norm, swish/sigmoid gate, residual, sorted-segment sum)
//   -> coalesced stores.
// Backward mirrors it: gate VJP -> H tiles -> gA = H W^T (row GEMMs), gW += A^T H (per-thread register
// accumulators that live across all tiles of the CTA, reduced deterministically by a second kernel),
// input-gradient assembly and scatter (store / red.v4 atomics / run-length-combined sorted adds).
#include <algorithm>
#include <cstdarg>
#include <vector>

#include "common.cuh"

namespace se3 {

static constexpr int NT = 256;  // threads per CTA
static constexpr float C3 = 0.57735026918962576451f;
static constexpr float C6 = 0.40824829046386301637f;

struct FamL {
    int ns, nd, nx, mz, mv;  // scalars, dot-vectors, cross-vectors ; l=0 outs, l=1 outs
    int nsp, ndp, nvp;       // pad4(ns), pad4(nd), pad4(nd+nx)
    int hasZ, hasV;
    int kz, kv;              // AZ / AV row strides (floats)
    int tnz, ncbz, tnv, ncbv;
    int t_s, t_d, t_x, t_oz, t_ov;  // column tables (offsets into tab)
    int nrm_z, nrm_v;               // offsets into the shared norm array
    int wz, wv;                     // species index of the weight matrices
    int o_az, o_av, o_wz, o_wvs, o_wvv;  // forward smem offsets (floats)
    // backward
    int mzp, mvp, mz4, mv4;
    int tn_s, ncb_s, tn_d, ncb_d, tn_t, ncb_t;
    int b_az, b_av, b_hz, b_hg, b_hv, b_wtzs, b_wtvs, b_wtzd, b_wtvv, b_gd, b_gt;
    int gdl, gtl;
    int nkt_zs, nkt_zd, nkt_vs, nkt_vv, nmt_z, nmt_v;
    int gw_z, gw_v;  // offsets into the flat weight-gradient buffer
    int job0, njobs; // gW thread-jobs
};

struct PlanL {
    int d_in1, d_out, TR, ntab, wtot, njw;
    int dop;  // out/post tile stride
    int gts;  // backward g tile stride
    FamL f[2];
    int o_y, o_rowoff, o_tab, o_norm, o_out, o_post, smem_fwd;
    int b_y, b_rowoff, b_tab, b_norm, b_gt, smem_bwd;
};

struct FwdK {
    long long rows;
    RowSrc src;
    const float* in2;
    const float* w[4];
    const float* norm[4];
    EpiL epi;
    float* out_raw;
    float* out_post;
    const float* resid;
    const int32_t* seg_idx;
    float* out_seg;
    const int* tab;
};

struct BwdK {
    long long rows;
    RowSrc src;
    const float* in2;
    const float* w[4];
    const float* norm[4];
    EpiL epi;
    const float* raw;
    const float* gout;
    const int32_t* gout_idx;
    float* gseg[SE3_MAX_SEG];
    int gmode[SE3_MAX_SEG];
    float* partials;  // [grid][wtot] or NULL
    const int* tab;
};

// ------------------------------------------------------------------ device helpers

__device__ __forceinline__ int slot_col(int slot, int tn) { return (slot >> 4) * 4 * tn + ((slot >> 2) & 3) * tn + (slot & 3); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

__device__ __forceinline__ const float* row_ptr(const RowSrc& rs, const long long* ro, int col) {
    int s = 0;
#pragma unroll
    for (int q = 1; q < SE3_MAX_SEG; ++q)
        if (q < rs.nseg && col >= rs.cum[q]) s = q;
    return rs.base[s] + ro[s] + (col - rs.cum[s]);
}

template <int TN>
__device__ __forceinline__ void gemm_rows(float (&acc)[4][TN], const float* __restrict__ a, int a_rstride, int k4_beg,
                                          int k4_end, const float* __restrict__ b, int ldb) {
#pragma unroll 2
    for (int k4 = k4_beg; k4 < k4_end; ++k4) {
        float4 av[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const float4*>(a + i * a_rstride + 4 * k4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const float4 w = *reinterpret_cast<const float4*>(b + (4 * k4 + kk) * ldb);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float x = f4c(av[i], kk);
                acc[i][0] = fmaf(x, w.x, acc[i][0]);
                if (TN > 1) acc[i][1 % TN] = fmaf(x, w.y, acc[i][1 % TN]);
                if (TN > 2) acc[i][2 % TN] = fmaf(x, w.z, acc[i][2 % TN]);
                if (TN > 3) acc[i][3 % TN] = fmaf(x, w.w, acc[i][3 % TN]);
            }
        }
    }
}

// Per-tile prologue shared by forward and backward: row offsets + in2 tile.
__device__ __forceinline__ void load_rows(const RowSrc& rs, const float* __restrict__ in2, long long row0, long long R,
                                          int TR, long long* rowoff, float* Yt) {
    for (int t = threadIdx.x; t < TR * SE3_MAX_SEG; t += NT) {
        const int row = t / SE3_MAX_SEG, s = t % SE3_MAX_SEG;
        const long long gr = row0 + row;
        long long off = 0;
        if (s < rs.nseg && gr < R) off = (long long)(rs.idx[s] ? rs.idx[s][gr] : gr) * rs.ld[s];
        rowoff[t] = off;
    }
    for (int t = threadIdx.x; t < TR * 4; t += NT) {
        const long long gr = row0 + (t >> 2);
        Yt[t] = gr < R ? __ldg(in2 + gr * 4 + (t & 3)) : 0.0f;
    }
}

// Feature tiles: AZ[row][0:nsp)=s, AZ[row][nsp:nsp+ndp)=c3<v,Y1>; AV[(row,c)][k]= c3*Y0*v_k[c] | c6*(x_k x Y1)[c]
__device__ __forceinline__ void build_features(const PlanL& P, const RowSrc& rs, float* sm, int off_az[2], int off_av[2],
                                               const int* tab, const long long* rowoff, const float* Yt, long long row0,
                                               long long R) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int row = warp; row < P.TR; row += NT / 32) {
        const bool valid = (row0 + row) < R;
        const long long* ro = rowoff + row * SE3_MAX_SEG;
        const float y0 = Yt[row * 4 + 0], y1x = Yt[row * 4 + 1], y1y = Yt[row * 4 + 2], y1z = Yt[row * 4 + 3];
#pragma unroll
        for (int fi = 0; fi < 2; ++fi) {
            const FamL& F = P.f[fi];
            if (!(F.hasZ | F.hasV)) continue;
            float* az = sm + off_az[fi] + row * F.kz;
            for (int k = lane; k < F.nsp; k += 32) {
                float v = 0.0f;
                if (valid && k < F.ns) v = __ldg(row_ptr(rs, ro, tab[F.t_s + k]));
                az[k] = v;
            }
            if (F.hasZ) {
                for (int k = lane; k < F.ndp; k += 32) {
                    float d = 0.0f;
                    if (valid && k < F.nd) {
                        const float* p = row_ptr(rs, ro, tab[F.t_d + k]);
                        d = C3 * (__ldg(p) * y1x + __ldg(p + 1) * y1y + __ldg(p + 2) * y1z);
                    }
                    az[F.nsp + k] = d;
                }
            }
            if (F.hasV) {
                float* av = sm + off_av[fi] + row * 3 * F.kv;
                for (int k = lane; k < F.nvp; k += 32) {
                    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
                    if (valid) {
                        if (k < F.nd) {
                            const float* p = row_ptr(rs, ro, tab[F.t_d + k]);
                            const float s = C3 * y0;
                            a0 = s * __ldg(p);
                            a1 = s * __ldg(p + 1);
                            a2 = s * __ldg(p + 2);
                        } else if (k < F.nd + F.nx) {
                            const float* p = row_ptr(rs, ro, tab[F.t_x + (k - F.nd)]);
                            const float vx = __ldg(p), vy = __ldg(p + 1), vz = __ldg(p + 2);
                            a0 = C6 * (vy * y1z - vz * y1y);
                            a1 = C6 * (vz * y1x - vx * y1z);
                            a2 = C6 * (vx * y1y - vy * y1x);
                        }
                    }
                    av[k] = a0;
                    av[F.kv + k] = a1;
                    av[2 * F.kv + k] = a2;
                }
            }
        }
    }
}

// ------------------------------------------------------------------ forward jobs

template <int TN>
__device__ __forceinline__ void job_z_fwd(const PlanL& P, const FamL& F, float* sm, const int* tab, const float* norm,
                                          int rgb, int cb, int lane) {
    const int rg = lane >> 2, cg = lane & 3;
    const int r0 = rgb * 32 + rg;
    const float* a = sm + F.o_az + r0 * F.kz;
    const int S = 16 * F.ncbz;
    const float* b = sm + F.o_wz + cb * 16 + cg * 4;
    float acc[4][TN];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;
    gemm_rows<TN>(acc, a, 8 * F.kz, 0, F.nsp >> 2, b, S);
    const float* Y = sm + P.o_y;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float y0 = Y[(r0 + 8 * i) * 4];
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] *= y0;
    }
    gemm_rows<TN>(acc, a, 8 * F.kz, F.nsp >> 2, (F.nsp + F.ndp) >> 2, b, S);
    float* out = sm + P.o_out;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        const int col = cb * 4 * TN + cg * TN + j;
        if (col < F.mz) {
            const int oc = tab[F.t_oz + col];
            const float nv = norm[F.nrm_z + col];
#pragma unroll
            for (int i = 0; i < 4; ++i) out[(r0 + 8 * i) * P.dop + oc] = acc[i][j] * nv;
        }
    }
}

template <int TN>
__device__ __forceinline__ void job_v_fwd(const PlanL& P, const FamL& F, float* sm, const int* tab, const float* norm,
                                          int rgb, int cb, int lane) {
    const int rg = lane >> 2, cg = lane & 3;
    const int r0 = rgb * 32 + rg;
    const int S = 16 * F.ncbv;
    float g[4][TN];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) g[i][j] = 0.0f;
    gemm_rows<TN>(g, sm + F.o_az + r0 * F.kz, 8 * F.kz, 0, F.nsp >> 2, sm + F.o_wvs + cb * 16 + cg * 4, S);
    const float* Y = sm + P.o_y;
    float* out = sm + P.o_out;
    const float* bv = sm + F.o_wvv + cb * 16 + cg * 4;
#pragma unroll 1
    for (int c = 0; c < 3; ++c) {
        float t[4][TN];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float s = C3 * Y[(r0 + 8 * i) * 4 + 1 + c];
#pragma unroll
            for (int j = 0; j < TN; ++j) t[i][j] = s * g[i][j];
        }
        gemm_rows<TN>(t, sm + F.o_av + (r0 * 3 + c) * F.kv, 24 * F.kv, 0, F.nvp >> 2, bv, S);
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int col = cb * 4 * TN + cg * TN + j;
            if (col < F.mv) {
                const int oc = tab[F.t_ov + col] + c;
                const float nv = norm[F.nrm_v + 3 * col + c];
#pragma unroll
                for (int i = 0; i < 4; ++i) out[(r0 + 8 * i) * P.dop + oc] = t[i][j] * nv;
            }
        }
    }
}

__device__ __forceinline__ void run_fwd_job(const PlanL& P, float* sm, const int* tab, const float* norm, int job,
                                            int lane) {
    const int nrg = P.TR >> 5;
#pragma unroll
    for (int fi = 0; fi < 2; ++fi) {
        const FamL& F = P.f[fi];
        const int nz = F.hasZ ? nrg * F.ncbz : 0;
        if (job < nz) {
            const int rgb = job / F.ncbz, cb = job % F.ncbz;
            switch (F.tnz) {
                case 1: job_z_fwd<1>(P, F, sm, tab, norm, rgb, cb, lane); break;
                case 2: job_z_fwd<2>(P, F, sm, tab, norm, rgb, cb, lane); break;
                case 3: job_z_fwd<3>(P, F, sm, tab, norm, rgb, cb, lane); break;
                default: job_z_fwd<4>(P, F, sm, tab, norm, rgb, cb, lane); break;
            }
            return;
        }
        job -= nz;
        const int nv = F.hasV ? nrg * F.ncbv : 0;
        if (job < nv) {
            const int rgb = job / F.ncbv, cb = job % F.ncbv;
            switch (F.tnv) {
                case 1: job_v_fwd<1>(P, F, sm, tab, norm, rgb, cb, lane); break;
                case 2: job_v_fwd<2>(P, F, sm, tab, norm, rgb, cb, lane); break;
                case 3: job_v_fwd<3>(P, F, sm, tab, norm, rgb, cb, lane); break;
                default: job_v_fwd<4>(P, F, sm, tab, norm, rgb, cb, lane); break;
            }
            return;
        }
        job -= nv;
    }
}

// weights -> shared memory in the slotted forward layouts
__device__ __forceinline__ void stage_weights_fwd(const PlanL& P, const FwdK& K, float* sm) {
#pragma unroll
    for (int fi = 0; fi < 2; ++fi) {
        const FamL& F = P.f[fi];
        if (F.hasZ) {
            const float* W = K.w[F.wz];
            const int S = 16 * F.ncbz, rows = F.nsp + F.ndp;
            for (int t = threadIdx.x; t < rows * S; t += NT) {
                const int kp = t / S, slot = t % S;
                const int col = slot_col(slot, F.tnz);
                int row = -1;
                if (kp < F.ns) row = kp;
                else if (kp >= F.nsp && kp < F.nsp + F.nd) row = F.ns + (kp - F.nsp);
                float v = 0.0f;
                if (row >= 0 && (slot & 3) < F.tnz && col < F.mz) v = __ldg(W + (long long)row * F.mz + col);
                sm[F.o_wz + t] = v;
            }
        }
        if (F.hasV) {
            const float* W = K.w[F.wv];
            const int S = 16 * F.ncbv;
            for (int t = threadIdx.x; t < F.nsp * S; t += NT) {
                const int kp = t / S, slot = t % S;
                const int col = slot_col(slot, F.tnv);
                float v = 0.0f;
                if (kp < F.ns && (slot & 3) < F.tnv && col < F.mv) v = __ldg(W + (long long)kp * F.mv + col);
                sm[F.o_wvs + t] = v;
            }
            for (int t = threadIdx.x; t < F.nvp * S; t += NT) {
                const int kp = t / S, slot = t % S;
                const int col = slot_col(slot, F.tnv);
                float v = 0.0f;
                if (kp < F.nd + F.nx && (slot & 3) < F.tnv && col < F.mv)
                    v = __ldg(W + (long long)(F.ns + kp) * F.mv + col);
                sm[F.o_wvv + t] = v;
            }
        }
    }
}

__device__ __forceinline__ void stage_tab_norm(const PlanL& P, const int* gtab, const float* const* gnorm, int* tab,
                                               float* norm) {
    for (int t = threadIdx.x; t < P.ntab; t += NT) tab[t] = gtab[t];
#pragma unroll
    for (int fi = 0; fi < 2; ++fi) {
        const FamL& F = P.f[fi];
        const float* nz = gnorm[F.wz];
        const float* nv = gnorm[F.wv];
        for (int t = threadIdx.x; t < F.mz; t += NT) norm[F.nrm_z + t] = nz ? nz[t] : 1.0f;
        for (int t = threadIdx.x; t < 3 * F.mv; t += NT) norm[F.nrm_v + t] = nv ? nv[t] : 1.0f;
    }
}

// sorted-segment sum of a [TR][width] shared tile into global rows keyed by `key[row]` (equal keys contiguous).
template <typename KeyT>
__device__ __forceinline__ void segsum_tile(const float* tile, int tstride, int col0, int width, const KeyT* key,
                                            int kstride, long long keymul, float* dst, int TR, int nvalid) {
    int parts = NT / width;
    if (parts < 1) parts = 1;
    if (parts > TR) parts = TR;
    const int rpp = (TR + parts - 1) / parts;
    for (int item = threadIdx.x; item < width * parts; item += NT) {
        const int c = item % width, q = item / width;
        const int rbeg = q * rpp;
        int rend = rbeg + rpp;
        if (rend > nvalid) rend = nvalid;
        if (rbeg >= rend) continue;
        KeyT cur = key[rbeg * kstride];
        float acc = 0.0f;
        for (int r = rbeg; r < rend; ++r) {
            const KeyT k = key[r * kstride];
            if (k != cur) {
                atomicAdd(dst + (long long)cur * keymul + c, acc);
                cur = k;
                acc = 0.0f;
            }
            acc += tile[r * tstride + col0 + c];
        }
        atomicAdd(dst + (long long)cur * keymul + c, acc);
    }
}

// ------------------------------------------------------------------ forward kernel

__global__ void __launch_bounds__(NT) l1tp_fwd_kernel(const PlanL P, const FwdK K) {
    extern __shared__ __align__(16) float sm[];
    float* Yt = sm + P.o_y;
    long long* rowoff = reinterpret_cast<long long*>(sm + P.o_rowoff);
    int* tab = reinterpret_cast<int*>(sm + P.o_tab);
    float* norm = sm + P.o_norm;
    float* out = sm + P.o_out;
    float* post = sm + P.o_post;
    int* segk = reinterpret_cast<int*>(sm + P.o_post + ((P.TR * P.dop + 3) & ~3));  // after the post tile

    stage_tab_norm(P, K.tab, K.norm, tab, norm);
    stage_weights_fwd(P, K, sm);
    __syncthreads();

    const int TR = P.TR;
    const long long R = K.rows;
    const long long ntiles = (R + TR - 1) / TR;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nrg = TR >> 5;
    int njobs = 0;
#pragma unroll
    for (int fi = 0; fi < 2; ++fi)
        njobs += (P.f[fi].hasZ ? nrg * P.f[fi].ncbz : 0) + (P.f[fi].hasV ? nrg * P.f[fi].ncbv : 0);
    int off_az[2] = {P.f[0].o_az, P.f[1].o_az};
    int off_av[2] = {P.f[0].o_av, P.f[1].o_av};

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long row0 = tile * TR;
        const int nvalid = (int)min((long long)TR, R - row0);
        load_rows(K.src, K.in2, row0, R, TR, rowoff, Yt);
        if (K.seg_idx)
            for (int t = threadIdx.x; t < TR; t += NT) segk[t] = t < nvalid ? K.seg_idx[row0 + t] : -1;
        __syncthreads();
        build_features(P, K.src, sm, off_az, off_av, tab, rowoff, Yt, row0, R);
        __syncthreads();
        for (int job = warp; job < njobs; job += NT / 32) run_fwd_job(P, sm, tab, norm, job, lane);
        __syncthreads();

        const int dout = P.d_out;
        if (K.out_raw) {
            float* dst = K.out_raw + row0 * dout;
            const float* res = K.resid ? K.resid + row0 * dout : nullptr;
            for (int t = threadIdx.x; t < nvalid * dout; t += NT) {
                const int row = t / dout, col = t - row * dout;
                float v = out[row * P.dop + col];
                if (res) v += __ldg(res + t);
                dst[t] = v;
            }
        }
        const float* stile = out;
        int swidth = dout;
        if (K.epi.mode == SE3_EPI_GATE) {
            const FamL& F = P.f[0];
            const int dp = K.epi.d_post, nsg = K.epi.ns_g;
            float* dst = K.out_post ? K.out_post + row0 * dp : nullptr;
            for (int t = threadIdx.x; t < nvalid * dp; t += NT) {
                const int row = t / dp, p = t - row * dp;
                const float* orow = out + row * P.dop;
                float v;
                if (p < nsg) {
                    const float x = orow[tab[F.t_oz + p]];
                    v = K.epi.cs * x * sigmoidf_(x);
                } else {
                    const int q = p - nsg;
                    const int vch = q / 3, c = q - vch * 3;
                    const float gx = orow[tab[F.t_oz + nsg + vch]];
                    v = K.epi.cg * sigmoidf_(gx) * orow[tab[F.t_ov + vch] + c];
                }
                post[row * P.dop + p] = v;
                if (dst) dst[t] = v;
            }
            stile = post;
            swidth = dp;
        } else if (K.resid && K.seg_idx) {
            // residual + segment-sum is not a supported combination (validated on the host)
        }
        if (K.seg_idx) {
            __syncthreads();
            segsum_tile<int>(stile, P.dop, 0, swidth, segk, 1, (long long)swidth, K.out_seg, TR, nvalid);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ backward

__device__ __forceinline__ void stage_weights_bwd(const PlanL& P, const BwdK& K, float* sm) {
#pragma unroll
    for (int fi = 0; fi < 2; ++fi) {
        const FamL& F = P.f[fi];
        const float* WZ = K.w[F.wz];
        const float* WV = K.w[F.wv];
        const int Ss = 16 * F.ncb_s, Sd = 16 * F.ncb_d, St = 16 * F.ncb_t;
        if (F.hasZ) {
            for (int t = threadIdx.x; t < F.mz4 * Ss; t += NT) {  // WtZs[m][slot(k<ns)]
                const int m = t / Ss, slot = t % Ss, k = slot_col(slot, F.tn_s);
                float v = 0.0f;
                if (m < F.mz && (slot & 3) < F.tn_s && k < F.ns) v = __ldg(WZ + (long long)k * F.mz + m);
                sm[F.b_wtzs + t] = v;
            }
            for (int t = threadIdx.x; t < F.mz4 * Sd; t += NT) {  // WtZd[m][slot(kd<nd)]
                const int m = t / Sd, slot = t % Sd, k = slot_col(slot, F.tn_d);
                float v = 0.0f;
                if (m < F.mz && (slot & 3) < F.tn_d && k < F.nd) v = __ldg(WZ + (long long)(F.ns + k) * F.mz + m);
                sm[F.b_wtzd + t] = v;
            }
        }
        if (F.hasV) {
            for (int t = threadIdx.x; t < F.mv4 * Ss; t += NT) {  // WtVs[m][slot(k<ns)]
                const int m = t / Ss, slot = t % Ss, k = slot_col(slot, F.tn_s);
                float v = 0.0f;
                if (m < F.mv && (slot & 3) < F.tn_s && k < F.ns) v = __ldg(WV + (long long)k * F.mv + m);
                sm[F.b_wtvs + t] = v;
            }
            for (int t = threadIdx.x; t < F.mv4 * St; t += NT) {  // WtVv[m][slot(kk<nd+nx)]
                const int m = t / St, slot = t % St, k = slot_col(slot, F.tn_t);
                float v = 0.0f;
                if (m < F.mv && (slot & 3) < F.tn_t && k < F.nd + F.nx) v = __ldg(WV + (long long)(F.ns + k) * F.mv + m);
                sm[F.b_wtvv + t] = v;
            }
        }
    }
}

// (i) gS[row][k<ns] = Y0 * HZ.WtZs + HG.WtVs  -> g tile, scalar columns
template <int TN>
__device__ __forceinline__ void job_gs(const PlanL& P, const FamL& F, float* sm, const int* tab, int rgb, int cb,
                                       int lane) {
    const int rg = lane >> 2, cg = lane & 3;
    const int r0 = rgb * 32 + rg;
    const int S = 16 * F.ncb_s;
    float acc[4][TN];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;
    if (F.hasZ) {
        gemm_rows<TN>(acc, sm + F.b_hz + r0 * F.mzp, 8 * F.mzp, 0, F.mz4 >> 2, sm + F.b_wtzs + cb * 16 + cg * 4, S);
        const float* Y = sm + P.b_y;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float y0 = Y[(r0 + 8 * i) * 4];
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] *= y0;
        }
    }
    if (F.hasV)
        gemm_rows<TN>(acc, sm + F.b_hg + r0 * F.mvp, 8 * F.mvp, 0, F.mv4 >> 2, sm + F.b_wtvs + cb * 16 + cg * 4, S);
    float* gt = sm + P.b_gt;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        const int k = cb * 4 * TN + cg * TN + j;
        if (k < F.ns) {
            const int col = tab[F.t_s + k];
#pragma unroll
            for (int i = 0; i < 4; ++i) gt[(r0 + 8 * i) * P.gts + col] = acc[i][j];
        }
    }
}

// (ii)/(iii) plain row GEMM: dst[row][k<n] = H[row][:] . Wt[:, slot(k)]
template <int TN>
__device__ __forceinline__ void job_plain(float* sm, int h_off, int ldh, int red4, int wt_off, int S, int dst_off,
                                          int ldd, int n, int rgb, int cb, int lane) {
    const int rg = lane >> 2, cg = lane & 3;
    const int r0 = rgb * 32 + rg;
    float acc[4][TN];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;
    gemm_rows<TN>(acc, sm + h_off + r0 * ldh, 8 * ldh, 0, red4, sm + wt_off + cb * 16 + cg * 4, S);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
        const int k = cb * 4 * TN + cg * TN + j;
        if (k < n) {
#pragma unroll
            for (int i = 0; i < 4; ++i) sm[dst_off + (r0 + 8 * i) * ldd + k] = acc[i][j];
        }
    }
}

__device__ __forceinline__ void run_plain(int tn, float* sm, int h_off, int ldh, int red4, int wt_off, int S,
                                          int dst_off, int ldd, int n, int rgb, int cb, int lane) {
    switch (tn) {
        case 1: job_plain<1>(sm, h_off, ldh, red4, wt_off, S, dst_off, ldd, n, rgb, cb, lane); break;
        case 2: job_plain<2>(sm, h_off, ldh, red4, wt_off, S, dst_off, ldd, n, rgb, cb, lane); break;
        case 3: job_plain<3>(sm, h_off, ldh, red4, wt_off, S, dst_off, ldd, n, rgb, cb, lane); break;
        default: job_plain<4>(sm, h_off, ldh, red4, wt_off, S, dst_off, ldd, n, rgb, cb, lane); break;
    }
}

__device__ __forceinline__ void run_bwd_job(const PlanL& P, float* sm, const int* tab, int job, int lane) {
    const int nrg = P.TR >> 5;
#pragma unroll
    for (int fi = 0; fi < 2; ++fi) {
        const FamL& F = P.f[fi];
        if (!(F.hasZ | F.hasV)) continue;
        const int n1 = F.ns > 0 ? nrg * F.ncb_s : 0;
        if (job < n1) {
            const int rgb = job / F.ncb_s, cb = job % F.ncb_s;
            switch (F.tn_s) {
                case 1: job_gs<1>(P, F, sm, tab, rgb, cb, lane); break;
                case 2: job_gs<2>(P, F, sm, tab, rgb, cb, lane); break;
                case 3: job_gs<3>(P, F, sm, tab, rgb, cb, lane); break;
                default: job_gs<4>(P, F, sm, tab, rgb, cb, lane); break;
            }
            return;
        }
        job -= n1;
        const int n2 = (F.hasZ && F.nd > 0) ? nrg * F.ncb_d : 0;
        if (job < n2) {
            run_plain(F.tn_d, sm, F.b_hz, F.mzp, F.mz4 >> 2, F.b_wtzd, 16 * F.ncb_d, F.b_gd, F.gdl, F.nd, job / F.ncb_d,
                      job % F.ncb_d, lane);
            return;
        }
        job -= n2;
        const int n3 = (F.hasV && F.nd + F.nx > 0) ? 3 * nrg * F.ncb_t : 0;
        if (job < n3) {
            run_plain(F.tn_t, sm, F.b_hv, F.mvp, F.mv4 >> 2, F.b_wtvv, 16 * F.ncb_t, F.b_gt, F.gtl, F.nd + F.nx,
                      job / F.ncb_t, job % F.ncb_t, lane);
            return;
        }
        job -= n3;
    }
}

__device__ __forceinline__ int count_bwd_jobs(const PlanL& P) {
    const int nrg = P.TR >> 5;
    int n = 0;
#pragma unroll
    for (int fi = 0; fi < 2; ++fi) {
        const FamL& F = P.f[fi];
        if (!(F.hasZ | F.hasV)) continue;
        n += F.ns > 0 ? nrg * F.ncb_s : 0;
        n += (F.hasZ && F.nd > 0) ? nrg * F.ncb_d : 0;
        n += (F.hasV && F.nd + F.nx > 0) ? 3 * nrg * F.ncb_t : 0;
    }
    return n;
}

// gW thread-job: 4 k-rows x 4 m-cols of one weight matrix, reduced over the rows of the tile.
struct WJob {
    int a_off, lda, h_off, ldh, nrows, scaled;
    int gw_off, krow0, kvalid, mcols, m0;  // output mapping
};

__device__ __forceinline__ bool decode_wjob(const PlanL& P, int j, WJob& J) {
#pragma unroll
    for (int fi = 0; fi < 2; ++fi) {
        const FamL& F = P.f[fi];
        if (j >= F.job0 && j < F.job0 + F.njobs) {
            int q = j - F.job0;
            const int n_zs = F.nkt_zs * F.nmt_z, n_zd = F.nkt_zd * F.nmt_z, n_vs = F.nkt_vs * F.nmt_v;
            if (q < n_zs) {
                const int kt = q / F.nmt_z, mt = q % F.nmt_z;
                J = {F.b_az + 4 * kt, F.kz, F.b_hz + 4 * mt, F.mzp, P.TR, 1, F.gw_z, 4 * kt, F.ns - 4 * kt, F.mz, 4 * mt};
                return true;
            }
            q -= n_zs;
            if (q < n_zd) {
                const int kt = q / F.nmt_z, mt = q % F.nmt_z;
                J = {F.b_az + F.nsp + 4 * kt, F.kz, F.b_hz + 4 * mt, F.mzp, P.TR, 0, F.gw_z, F.ns + 4 * kt, F.nd - 4 * kt,
                     F.mz, 4 * mt};
                return true;
            }
            q -= n_zd;
            if (q < n_vs) {
                const int kt = q / F.nmt_v, mt = q % F.nmt_v;
                J = {F.b_az + 4 * kt, F.kz, F.b_hg + 4 * mt, F.mvp, P.TR, 0, F.gw_v, 4 * kt, F.ns - 4 * kt, F.mv, 4 * mt};
                return true;
            }
            q -= n_vs;
            const int kt = q / F.nmt_v, mt = q % F.nmt_v;
            J = {F.b_av + 4 * kt, F.kv, F.b_hv + 4 * mt, F.mvp, 3 * P.TR, 0, F.gw_v, F.ns + 4 * kt, F.nd + F.nx - 4 * kt,
                 F.mv, 4 * mt};
            return true;
        }
    }
    return false;
}

__device__ __forceinline__ void wjob_accumulate(const WJob& J, const float* sm, const float* Yt, float (&acc)[16]) {
    const float* a = sm + J.a_off;
    const float* h = sm + J.h_off;
#pragma unroll 4
    for (int r = 0; r < J.nrows; ++r) {
        const float4 av = *reinterpret_cast<const float4*>(a + r * J.lda);
        float4 hv = *reinterpret_cast<const float4*>(h + r * J.ldh);
        if (J.scaled) {
            const float y0 = Yt[r * 4];
            hv.x *= y0; hv.y *= y0; hv.z *= y0; hv.w *= y0;
        }
        acc[0] = fmaf(av.x, hv.x, acc[0]);   acc[1] = fmaf(av.x, hv.y, acc[1]);
        acc[2] = fmaf(av.x, hv.z, acc[2]);   acc[3] = fmaf(av.x, hv.w, acc[3]);
        acc[4] = fmaf(av.y, hv.x, acc[4]);   acc[5] = fmaf(av.y, hv.y, acc[5]);
        acc[6] = fmaf(av.y, hv.z, acc[6]);   acc[7] = fmaf(av.y, hv.w, acc[7]);
        acc[8] = fmaf(av.z, hv.x, acc[8]);   acc[9] = fmaf(av.z, hv.y, acc[9]);
        acc[10] = fmaf(av.z, hv.z, acc[10]); acc[11] = fmaf(av.z, hv.w, acc[11]);
        acc[12] = fmaf(av.w, hv.x, acc[12]); acc[13] = fmaf(av.w, hv.y, acc[13]);
        acc[14] = fmaf(av.w, hv.z, acc[14]); acc[15] = fmaf(av.w, hv.w, acc[15]);
    }
}

template <int MAXWJ>
__global__ void __launch_bounds__(NT) l1tp_bwd_kernel(const PlanL P, const BwdK K) {
    extern __shared__ __align__(16) float sm[];
    float* Yt = sm + P.b_y;
    long long* rowoff = reinterpret_cast<long long*>(sm + P.b_rowoff);
    int* tab = reinterpret_cast<int*>(sm + P.b_tab);
    float* norm = sm + P.b_norm;
    float* gt = sm + P.b_gt;

    stage_tab_norm(P, K.tab, K.norm, tab, norm);
    stage_weights_bwd(P, K, sm);
    __syncthreads();

    const int TR = P.TR;
    const long long R = K.rows;
    const long long ntiles = (R + TR - 1) / TR;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int njobs = count_bwd_jobs(P);
    int off_az[2] = {P.f[0].b_az, P.f[1].b_az};
    int off_av[2] = {P.f[0].b_av, P.f[1].b_av};
    const bool do_gw = K.partials != nullptr;

    float wacc[MAXWJ][16];
#pragma unroll
    for (int j = 0; j < MAXWJ; ++j)
#pragma unroll
        for (int e = 0; e < 16; ++e) wacc[j][e] = 0.0f;

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long row0 = tile * TR;
        const int nvalid = (int)min((long long)TR, R - row0);
        load_rows(K.src, K.in2, row0, R, TR, rowoff, Yt);
        __syncthreads();
        if (do_gw) build_features(P, K.src, sm, off_az, off_av, tab, rowoff, Yt, row0, R);  // only gW needs A tiles

        // ---- cotangent of the raw TP output -> g tile (aliases the input-gradient tile)
        const int dout = P.d_out;
        if (K.epi.mode == SE3_EPI_GATE) {
            const FamL& F = P.f[0];
            const int nsg = K.epi.ns_g, items = K.epi.ns_g + K.epi.nv, dp = K.epi.d_post;
            for (int t = threadIdx.x; t < TR * items; t += NT) {
                const int row = t / items, it = t - row * items;
                float* grow = gt + row * P.gts;
                if (row >= nvalid) {
                    if (it < nsg) grow[tab[F.t_oz + it]] = 0.0f;
                    else {
                        const int v = it - nsg, vc = tab[F.t_ov + v];
                        grow[tab[F.t_oz + nsg + v]] = 0.0f;
                        grow[vc] = 0.0f; grow[vc + 1] = 0.0f; grow[vc + 2] = 0.0f;
                    }
                    continue;
                }
                const long long gr = row0 + row;
                const float* rawr = K.raw + gr * dout;
                const float* gor = K.gout + (long long)(K.gout_idx ? K.gout_idx[gr] : gr) * dp;
                if (it < nsg) {
                    const int rc = tab[F.t_oz + it];
                    const float x = __ldg(rawr + rc), s = sigmoidf_(x);
                    grow[rc] = __ldg(gor + it) * K.epi.cs * s * (1.0f + x * (1.0f - s));
                } else {
                    const int v = it - nsg;
                    const int gc = tab[F.t_oz + nsg + v], vc = tab[F.t_ov + v];
                    const float s = sigmoidf_(__ldg(rawr + gc));
                    const float g0 = __ldg(gor + nsg + 3 * v), g1 = __ldg(gor + nsg + 3 * v + 1),
                                g2 = __ldg(gor + nsg + 3 * v + 2);
                    const float r0 = __ldg(rawr + vc), r1 = __ldg(rawr + vc + 1), r2 = __ldg(rawr + vc + 2);
                    grow[gc] = K.epi.cg * s * (1.0f - s) * (g0 * r0 + g1 * r1 + g2 * r2);
                    const float cs_ = K.epi.cg * s;
                    grow[vc] = cs_ * g0; grow[vc + 1] = cs_ * g1; grow[vc + 2] = cs_ * g2;
                }
            }
        } else {
            for (int t = threadIdx.x; t < TR * dout; t += NT) {
                const int row = t / dout, col = t - row * dout;
                float v = 0.0f;
                if (row < nvalid) {
                    const long long gr = row0 + row;
                    v = __ldg(K.gout + (long long)(K.gout_idx ? K.gout_idx[gr] : gr) * dout + col);
                }
                gt[row * P.gts + col] = v;
            }
        }
        __syncthreads();
        // ---- H tiles
#pragma unroll
        for (int fi = 0; fi < 2; ++fi) {
            const FamL& F = P.f[fi];
            if (F.hasZ) {
                for (int t = threadIdx.x; t < TR * F.mzp; t += NT) {
                    const int row = t / F.mzp, m = t - row * F.mzp;
                    sm[F.b_hz + t] = m < F.mz ? norm[F.nrm_z + m] * gt[row * P.gts + tab[F.t_oz + m]] : 0.0f;
                }
            }
            if (F.hasV) {
                for (int t = threadIdx.x; t < TR * F.mvp; t += NT) {
                    const int row = t / F.mvp, m = t - row * F.mvp;
                    float h0 = 0.0f, h1 = 0.0f, h2 = 0.0f;
                    if (m < F.mv) {
                        const float* g = gt + row * P.gts + tab[F.t_ov + m];
                        h0 = norm[F.nrm_v + 3 * m] * g[0];
                        h1 = norm[F.nrm_v + 3 * m + 1] * g[1];
                        h2 = norm[F.nrm_v + 3 * m + 2] * g[2];
                    }
                    float* hv = sm + F.b_hv + row * 3 * F.mvp + m;
                    hv[0] = h0; hv[F.mvp] = h1; hv[2 * F.mvp] = h2;
                    sm[F.b_hg + t] = C3 * (Yt[row * 4 + 1] * h0 + Yt[row * 4 + 2] * h1 + Yt[row * 4 + 3] * h2);
                }
            }
        }
        __syncthreads();
        // ---- weight-gradient thread jobs (register accumulators persist across tiles)
        if (do_gw) {
#pragma unroll
            for (int j = 0; j < MAXWJ; ++j) {
                const int jj = threadIdx.x + j * NT;
                WJob J;
                if (jj < P.njw && decode_wjob(P, jj, J)) wjob_accumulate(J, sm, Yt, wacc[j]);
            }
        }
        // ---- input-gradient row GEMMs
        for (int job = warp; job < njobs; job += NT / 32) run_bwd_job(P, sm, tab, job, lane);
        __syncthreads();
        // ---- assemble vector-input gradients (and zero the columns no path touches)
#pragma unroll
        for (int fi = 0; fi < 2; ++fi) {
            const FamL& F = P.f[fi];
            const FamL& G = P.f[1 - fi];
            const bool factive = F.hasZ | F.hasV;
            if (!factive)
                for (int t = threadIdx.x; t < TR * F.ns; t += NT) {
                    const int row = t / F.ns, k = t - row * F.ns;
                    gt[row * P.gts + tab[F.t_s + k]] = 0.0f;
                }
            for (int t = threadIdx.x; t < TR * F.nd; t += NT) {
                const int row = t / F.nd, kd = t - row * F.nd;
                const float y0 = Yt[row * 4], y1x = Yt[row * 4 + 1], y1y = Yt[row * 4 + 2], y1z = Yt[row * 4 + 3];
                float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f;
                if (F.hasZ) {
                    const float gd = C3 * sm[F.b_gd + row * F.gdl + kd];
                    g0 = gd * y1x; g1 = gd * y1y; g2 = gd * y1z;
                }
                if (F.hasV) {
                    const float* q = sm + F.b_gt + row * 3 * F.gtl + kd;
                    const float s = C3 * y0;
                    g0 += s * q[0]; g1 += s * q[F.gtl]; g2 += s * q[2 * F.gtl];
                }
                if (G.hasV) {  // this species is G's cross input: d/dv of gx.(v x Y1) = Y1 x gx
                    const float* q = sm + G.b_gt + row * 3 * G.gtl + G.nd + kd;
                    const float x0 = q[0], x1 = q[G.gtl], x2 = q[2 * G.gtl];
                    g0 += C6 * (y1y * x2 - y1z * x1);
                    g1 += C6 * (y1z * x0 - y1x * x2);
                    g2 += C6 * (y1x * x1 - y1y * x0);
                }
                float* d = gt + row * P.gts + tab[F.t_d + kd];
                d[0] = g0; d[1] = g1; d[2] = g2;
            }
        }
        __syncthreads();
        // ---- scatter the input-gradient tile
        for (int s = 0; s < K.src.nseg; ++s) {
            float* gb = K.gseg[s];
            const int mode = K.gmode[s];
            if (!gb || mode == SE3_GRAD_NONE) continue;
            const int w = K.src.cum[s + 1] - K.src.cum[s], c0 = K.src.cum[s];
            if (mode == SE3_GRAD_STORE) {
                for (int t = threadIdx.x; t < nvalid * w; t += NT) {
                    const int row = t / w, c = t - row * w;
                    gb[rowoff[row * SE3_MAX_SEG + s] + c] = gt[row * P.gts + c0 + c];
                }
            } else if (mode == SE3_GRAD_ATOMIC) {
                if ((w & 3) == 0 && (K.src.ld[s] & 3) == 0 && ((uintptr_t)gb & 15) == 0) {
                    const int w4 = w >> 2;
                    for (int t = threadIdx.x; t < nvalid * w4; t += NT) {
                        const int row = t / w4, c = (t - row * w4) << 2;
                        const float* g = gt + row * P.gts + c0 + c;
                        red_add_v4(gb + rowoff[row * SE3_MAX_SEG + s] + c, g[0], g[1], g[2], g[3]);
                    }
                } else {
                    for (int t = threadIdx.x; t < nvalid * w; t += NT) {
                        const int row = t / w, c = t - row * w;
                        atomicAdd(gb + rowoff[row * SE3_MAX_SEG + s] + c, gt[row * P.gts + c0 + c]);
                    }
                }
            } else {  // SE3_GRAD_SORTED: key = element offset of the destination row
                segsum_tile<long long>(gt, P.gts, c0, w, rowoff + s, SE3_MAX_SEG, 1LL, gb, TR, nvalid);
            }
        }
        __syncthreads();
    }

    if (do_gw) {
        float* part = K.partials + (long long)blockIdx.x * P.wtot;
#pragma unroll
        for (int j = 0; j < MAXWJ; ++j) {
            const int jj = threadIdx.x + j * NT;
            WJob J;
            if (jj < P.njw && decode_wjob(P, jj, J)) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (i < J.kvalid && J.m0 + e < J.mcols)
                            part[J.gw_off + (long long)(J.krow0 + i) * J.mcols + J.m0 + e] = wacc[j][i * 4 + e];
            }
        }
    }
}

struct ReduceK {
    const float* partials;
    int nblocks, wtot;
    float* gw[4];
    int off[4], cnt[4];
};

__global__ void l1tp_reduce_gw_kernel(const ReduceK K) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= K.wtot) return;
    float s = 0.0f;
    for (int b = 0; b < K.nblocks; ++b) s += K.partials[(long long)b * K.wtot + e];
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (K.gw[q] && e >= K.off[q] && e < K.off[q] + K.cnt[q]) K.gw[q][e - K.off[q]] = s;
}

// Gradient w.r.t. in2 (only the drop-in module needs it; RAW epilogue).  One warp per row, lanes over
// output channels; weights are read through L1/L2 (they are a few KB).
struct GIn2K {
    long long rows;
    RowSrc src;
    const float* in2;
    const float* w[4];
    const float* norm[4];
    const float* gout;
    float* gin2;
    const int* tab;
};

__global__ void __launch_bounds__(NT) l1tp_gin2_kernel(const PlanL P, const GIn2K K) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
    if (row >= K.rows) return;
    long long ro[SE3_MAX_SEG];
#pragma unroll
    for (int s = 0; s < SE3_MAX_SEG; ++s)
        ro[s] = s < K.src.nseg ? (long long)(K.src.idx[s] ? K.src.idx[s][row] : row) * K.src.ld[s] : 0;
    const float* go = K.gout + row * P.d_out;
    float gy0 = 0.0f, gyx = 0.0f, gyy = 0.0f, gyz = 0.0f;
    const int* tab = K.tab;
#pragma unroll
    for (int fi = 0; fi < 2; ++fi) {
        const FamL& F = P.f[fi];
        if (F.hasZ) {
            const float* W = K.w[F.wz];
            const float* nz = K.norm[F.wz];
            for (int m = lane; m < F.mz; m += 32) {
                const float h = (nz ? nz[m] : 1.0f) * go[tab[F.t_oz + m]];
                float u = 0.0f, ux = 0.0f, uy = 0.0f, uz = 0.0f;
                for (int k = 0; k < F.ns; ++k) u = fmaf(__ldg(row_ptr(K.src, ro, tab[F.t_s + k])), W[k * F.mz + m], u);
                for (int k = 0; k < F.nd; ++k) {
                    const float* p = row_ptr(K.src, ro, tab[F.t_d + k]);
                    const float w = W[(F.ns + k) * F.mz + m];
                    ux = fmaf(__ldg(p), w, ux); uy = fmaf(__ldg(p + 1), w, uy); uz = fmaf(__ldg(p + 2), w, uz);
                }
                gy0 += h * u;
                gyx += C3 * h * ux; gyy += C3 * h * uy; gyz += C3 * h * uz;
            }
        }
        if (F.hasV) {
            const float* W = K.w[F.wv];
            const float* nv = K.norm[F.wv];
            for (int m = lane; m < F.mv; m += 32) {
                const int oc = tab[F.t_ov + m];
                const float hx = (nv ? nv[3 * m] : 1.0f) * go[oc], hy = (nv ? nv[3 * m + 1] : 1.0f) * go[oc + 1],
                            hz = (nv ? nv[3 * m + 2] : 1.0f) * go[oc + 2];
                float g = 0.0f, vx = 0.0f, vy = 0.0f, vz = 0.0f, zx = 0.0f, zy = 0.0f, zz = 0.0f;
                for (int k = 0; k < F.ns; ++k) g = fmaf(__ldg(row_ptr(K.src, ro, tab[F.t_s + k])), W[k * F.mv + m], g);
                for (int k = 0; k < F.nd; ++k) {
                    const float* p = row_ptr(K.src, ro, tab[F.t_d + k]);
                    const float w = W[(F.ns + k) * F.mv + m];
                    vx = fmaf(__ldg(p), w, vx); vy = fmaf(__ldg(p + 1), w, vy); vz = fmaf(__ldg(p + 2), w, vz);
                }
                for (int k = 0; k < F.nx; ++k) {
                    const float* p = row_ptr(K.src, ro, tab[F.t_x + k]);
                    const float w = W[(F.ns + F.nd + k) * F.mv + m];
                    zx = fmaf(__ldg(p), w, zx); zy = fmaf(__ldg(p + 1), w, zy); zz = fmaf(__ldg(p + 2), w, zz);
                }
                gyx += C3 * hx * g; gyy += C3 * hy * g; gyz += C3 * hz * g;
                gy0 += C3 * (hx * vx + hy * vy + hz * vz);
                // out += c6 (Z x Y1)  =>  gY1 += c6 (h x Z)
                gyx += C6 * (hy * zz - hz * zy);
                gyy += C6 * (hz * zx - hx * zz);
                gyz += C6 * (hx * zy - hy * zx);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        gy0 += __shfl_xor_sync(0xffffffffu, gy0, o);
        gyx += __shfl_xor_sync(0xffffffffu, gyx, o);
        gyy += __shfl_xor_sync(0xffffffffu, gyy, o);
        gyz += __shfl_xor_sync(0xffffffffu, gyz, o);
    }
    if (lane == 0) {
        float4 r = make_float4(gy0, gyx, gyy, gyz);
        *reinterpret_cast<float4*>(K.gin2 + row * 4) = r;
    }
}

// ------------------------------------------------------------------ host: layout

static void slots(int n, int& tn, int& ncb) {
    ncb = (n + 15) / 16;
    if (ncb < 1) ncb = 1;
    tn = (n + 4 * ncb - 1) / (4 * ncb);
    if (tn < 1) tn = 1;
}

static void compute_layout(PlanL& L, int TR) {
    L.TR = TR;
    L.dop = L.d_out | 1;
    L.gts = std::max(L.d_in1, L.d_out) | 1;
    auto A4 = [](int x) { return (x + 3) & ~3; };
    // ---- forward
    int o = 0;
    L.o_y = o; o += TR * 4;
    L.o_rowoff = o; o += TR * SE3_MAX_SEG * 2;
    L.o_tab = o; o += A4(L.ntab);
    L.o_norm = o; o += A4(L.d_out);
    for (int fi = 0; fi < 2; ++fi) {
        FamL& F = L.f[fi];
        F.o_az = F.o_av = F.o_wz = F.o_wvs = F.o_wvv = 0;
        if (!(F.hasZ || F.hasV)) continue;
        F.o_az = o; o += TR * F.kz;
        if (F.hasV) { F.o_av = o; o += 3 * TR * F.kv; }
        if (F.hasZ) { F.o_wz = o; o += (F.nsp + F.ndp) * 16 * F.ncbz; }
        if (F.hasV) {
            F.o_wvs = o; o += F.nsp * 16 * F.ncbv;
            F.o_wvv = o; o += F.nvp * 16 * F.ncbv;
        }
    }
    L.o_out = o; o += A4(TR * L.dop);
    L.o_post = o; o += A4(TR * L.dop) + A4(TR);
    L.smem_fwd = o * 4;
    // ---- backward
    o = 0;
    L.b_y = o; o += TR * 4;
    L.b_rowoff = o; o += TR * SE3_MAX_SEG * 2;
    L.b_tab = o; o += A4(L.ntab);
    L.b_norm = o; o += A4(L.d_out);
    for (int fi = 0; fi < 2; ++fi) {
        FamL& F = L.f[fi];
        F.b_az = F.b_av = F.b_hz = F.b_hg = F.b_hv = F.b_wtzs = F.b_wtvs = F.b_wtzd = F.b_wtvv = F.b_gd = F.b_gt = 0;
        if (!(F.hasZ || F.hasV)) continue;
        F.b_az = o; o += TR * F.kz;
        if (F.hasV) { F.b_av = o; o += 3 * TR * F.kv; }
        if (F.hasZ) {
            F.b_hz = o; o += TR * F.mzp;
            F.b_wtzs = o; o += F.mz4 * 16 * F.ncb_s;
            F.b_wtzd = o; o += F.mz4 * 16 * F.ncb_d;
            F.b_gd = o; o += A4(TR * F.gdl);
        }
        if (F.hasV) {
            F.b_hg = o; o += TR * F.mvp;
            F.b_hv = o; o += 3 * TR * F.mvp;
            F.b_wtvs = o; o += F.mv4 * 16 * F.ncb_s;
            F.b_wtvv = o; o += F.mv4 * 16 * F.ncb_t;
            F.b_gt = o; o += A4(3 * TR * F.gtl);
        }
    }
    L.b_gt = o; o += A4(TR * L.gts);
    L.smem_bwd = o * 4;
}

}  // namespace se3

using namespace se3;

struct se3_l1tp_plan {
    PlanL L;
    int* d_tab = nullptr;
    float* d_partials = nullptr;
    size_t partial_cap = 0;  // floats
    int occ_fwd = 1, occ_bwd = 1, maxwj = 2;
    int w_off[4], w_cnt[4];
    int n[4], m[4];
    int t_in[4], t_out[4];
    std::vector<int> h_tab;   // host copy of the column tables (slot assignment of the tcgen05 kernels)
};

int se3_l1tp_tc2_try_forward(const int n[4], const int m[4], const int t_in[4], const int t_out[4], int ntab,
                             const int* h_tab, const int* d_tab, const se3_l1tp_fwd_args* a, const se3::RowSrc& src,
                             const se3::EpiL& epi, cudaStream_t st, bool* launched);
int se3_l1tp_tc2_try_backward_in(const int n[4], const int m[4], const int t_in[4], const int t_out[4], int ntab,
                                 const int* h_tab, const int* d_tab, const se3_l1tp_bwd_args* a, const se3::RowSrc& src,
                                 const se3::EpiL& epi, float* const gseg[SE3_MAX_SEG], const int gmode[SE3_MAX_SEG],
                                 cudaStream_t st, bool* launched);
int se3_l1tp_tc2_try_backward_w(const int n[4], const int m[4], const int t_in[4], const int t_out[4], int ntab,
                                const int* h_tab, const se3_l1tp_bwd_args* a, const se3::RowSrc& src, const se3::EpiL& epi,
                                float* partials, int wtot, int gw_z_off, int gw_v_off, int max_grid, cudaStream_t st,
                                int* grid_out, bool* launched);
extern "C" int se3_l1tp_plan_create(const se3_l1tp_desc* d, se3_l1tp_plan** out) {
    if (!d || !out) { set_error("null argument"); return SE3_ERR_INVALID; }
    *out = nullptr;
    int dev = 0;
    SE3_CUDA_TRY(cudaGetDevice(&dev));
    int din = d->n[0] + d->n[1] + 3 * (d->n[2] + d->n[3]);
    int dout = d->m[0] + d->m[1] + 3 * (d->m[2] + d->m[3]);
    if (din != d->d_in1 || dout != d->d_out || din <= 0 || dout <= 0) {
        set_error("species counts do not add up to d_in1/d_out");
        return SE3_ERR_INVALID;
    }
    se3_l1tp_plan* p = new se3_l1tp_plan();
    PlanL& L = p->L;
    memset(&L, 0, sizeof(L));
    L.d_in1 = din;
    L.d_out = dout;
    // column tables: in sp0..3, out sp0..3
    std::vector<int> tab;
    int t_in[4], t_out[4];
    for (int s = 0; s < 4; ++s) {
        t_in[s] = (int)tab.size();
        for (int k = 0; k < d->n[s]; ++k) tab.push_back(d->in_cols[s][k]);
    }
    for (int s = 0; s < 4; ++s) {
        t_out[s] = (int)tab.size();
        for (int k = 0; k < d->m[s]; ++k) tab.push_back(d->out_cols[s][k]);
    }
    for (int s = 0; s < 4; ++s) { p->n[s] = d->n[s]; p->m[s] = d->m[s]; p->t_in[s] = t_in[s]; p->t_out[s] = t_out[s]; }
    L.ntab = (int)tab.size();
    p->h_tab = tab;
    // families: E = (s 0e, dot 1o, cross 1e -> Z 0e, V 1o), O = (s 0o, dot 1e, cross 1o -> Z 0o, V 1e)
    const int fs[2] = {0, 1}, fd[2] = {3, 2}, fx[2] = {2, 3}, fz[2] = {0, 1}, fv[2] = {3, 2};
    int wtot = 0, njw = 0, nrm = 0;
    for (int s = 0; s < 4; ++s) { p->w_off[s] = 0; p->w_cnt[s] = 0; }
    for (int fi = 0; fi < 2; ++fi) {
        FamL& F = L.f[fi];
        F.ns = d->n[fs[fi]]; F.nd = d->n[fd[fi]]; F.nx = d->n[fx[fi]];
        F.mz = d->m[fz[fi]]; F.mv = d->m[fv[fi]];
        F.wz = fz[fi]; F.wv = fv[fi];
        F.t_s = t_in[fs[fi]]; F.t_d = t_in[fd[fi]]; F.t_x = t_in[fx[fi]];
        F.t_oz = t_out[fz[fi]]; F.t_ov = t_out[fv[fi]];
        F.hasZ = (F.mz > 0 && F.ns + F.nd > 0) ? 1 : 0;
        F.hasV = (F.mv > 0 && F.ns + F.nd + F.nx > 0) ? 1 : 0;
        if ((F.mz > 0 && !F.hasZ) || (F.mv > 0 && !F.hasV)) {
            set_error("output species without any contributing input (the reference cannot construct this either)");
            delete p;
            return SE3_ERR_INVALID;
        }
        F.nsp = pad4(F.ns); F.ndp = pad4(F.nd); F.nvp = pad4(F.nd + F.nx);
        F.kz = stride4odd(F.nsp + (F.hasZ ? F.ndp : 0));
        F.kv = stride4odd(F.nvp);
        slots(F.mz, F.tnz, F.ncbz);
        slots(F.mv, F.tnv, F.ncbv);
        F.nrm_z = nrm; nrm += F.mz;
        F.nrm_v = nrm; nrm += 3 * F.mv;
        F.mz4 = pad4(F.mz); F.mv4 = pad4(F.mv);
        F.mzp = stride4odd(F.mz4); F.mvp = stride4odd(F.mv4);
        slots(F.ns, F.tn_s, F.ncb_s);
        slots(F.nd, F.tn_d, F.ncb_d);
        slots(F.nd + F.nx, F.tn_t, F.ncb_t);
        F.gdl = F.nd | 1;
        F.gtl = (F.nd + F.nx) | 1;
        F.nkt_zs = F.hasZ ? F.nsp / 4 : 0;
        F.nkt_zd = F.hasZ ? F.ndp / 4 : 0;
        F.nkt_vs = F.hasV ? F.nsp / 4 : 0;
        F.nkt_vv = F.hasV ? F.nvp / 4 : 0;
        F.nmt_z = F.mz4 / 4; F.nmt_v = F.mv4 / 4;
        F.job0 = njw;
        F.njobs = (F.nkt_zs + F.nkt_zd) * F.nmt_z + (F.nkt_vs + F.nkt_vv) * F.nmt_v;
        njw += F.njobs;
        F.gw_z = wtot;
        if (F.hasZ) { p->w_off[F.wz] = wtot; p->w_cnt[F.wz] = (F.ns + F.nd) * F.mz; wtot += p->w_cnt[F.wz]; }
        F.gw_v = wtot;
        if (F.hasV) { p->w_off[F.wv] = wtot; p->w_cnt[F.wv] = (F.ns + F.nd + F.nx) * F.mv; wtot += p->w_cnt[F.wv]; }
    }
    L.wtot = wtot;
    L.njw = njw;
    if (njw > 4 * NT) {
        set_error("weight matrices too large for the register-resident gradient tiling (%d > %d 4x4 tiles)", njw, 4 * NT);
        delete p;
        return SE3_ERR_TOO_LARGE;
    }
    p->maxwj = njw <= 2 * NT ? 2 : 4;
    // pick the largest row tile that fits; prefer two resident CTAs per SM
    int maxsm = 0;
    SE3_CUDA_TRY(cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int two_cta = (228 * 1024) / 2 - 1024 - 512;
    int chosen = 0;
    for (int pass = 0; pass < 2 && !chosen; ++pass) {
        const int limit = pass == 0 ? two_cta : maxsm;
        for (int TR : {64, 32}) {
            compute_layout(L, TR);
            if (std::max(L.smem_fwd, L.smem_bwd) <= limit) { chosen = TR; break; }
        }
    }
    if (!chosen) {
        set_error("irreps too large for shared memory (need %d B fwd / %d B bwd at 32 rows, have %d)", L.smem_fwd,
                  L.smem_bwd, maxsm);
        delete p;
        return SE3_ERR_TOO_LARGE;
    }
    SE3_CUDA_TRY(cudaMalloc(&p->d_tab, sizeof(int) * std::max(1, L.ntab)));
    SE3_CUDA_TRY(cudaMemcpy(p->d_tab, tab.data(), sizeof(int) * L.ntab, cudaMemcpyHostToDevice));
    SE3_CUDA_TRY(cudaFuncSetAttribute(l1tp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
    SE3_CUDA_TRY(cudaFuncSetAttribute(l1tp_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
    SE3_CUDA_TRY(cudaFuncSetAttribute(l1tp_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
    SE3_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p->occ_fwd, l1tp_fwd_kernel, NT, L.smem_fwd));
    if (p->maxwj == 2)
        SE3_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p->occ_bwd, l1tp_bwd_kernel<2>, NT, L.smem_bwd));
    else
        SE3_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p->occ_bwd, l1tp_bwd_kernel<4>, NT, L.smem_bwd));
    if (p->occ_fwd < 1 || p->occ_bwd < 1) {
        set_error("kernel does not fit on an SM (occupancy 0)");
        se3_l1tp_plan_destroy(p);
        return SE3_ERR_TOO_LARGE;
    }
    *out = p;
    return SE3_OK;
}

extern "C" void se3_l1tp_plan_destroy(se3_l1tp_plan* p) {
    if (!p) return;
    if (p->d_tab) cudaFree(p->d_tab);
    if (p->d_partials) cudaFree(p->d_partials);
    delete p;
}

extern "C" int se3_l1tp_plan_info(const se3_l1tp_plan* p, int32_t* tile_rows, int32_t* smem_fwd, int32_t* smem_bwd,
                                  int32_t* weight_floats) {
    if (!p) return SE3_ERR_INVALID;
    if (tile_rows) *tile_rows = p->L.TR;
    if (smem_fwd) *smem_fwd = p->L.smem_fwd;
    if (smem_bwd) *smem_bwd = p->L.smem_bwd;
    if (weight_floats) *weight_floats = p->L.wtot;
    return SE3_OK;
}

static int fill_src(const se3_l1tp_plan* p, int nseg, const se3_rowseg* seg, RowSrc& rs) {
    if (nseg < 1 || nseg > SE3_MAX_SEG) { set_error("nseg must be 1..%d", SE3_MAX_SEG); return SE3_ERR_INVALID; }
    memset(&rs, 0, sizeof(rs));
    rs.nseg = nseg;
    int cum = 0;
    for (int s = 0; s < nseg; ++s) {
        if (!seg[s].base || seg[s].width <= 0 || seg[s].ld < seg[s].width) {
            set_error("segment %d: bad base/width/ld", s);
            return SE3_ERR_INVALID;
        }
        rs.base[s] = seg[s].base; rs.idx[s] = seg[s].idx; rs.ld[s] = seg[s].ld; rs.cum[s] = cum;
        cum += seg[s].width;
    }
    for (int s = nseg; s <= SE3_MAX_SEG; ++s) rs.cum[s] = cum;
    if (cum != p->L.d_in1) { set_error("segment widths sum to %d, in1 width is %d", cum, p->L.d_in1); return SE3_ERR_INVALID; }
    return SE3_OK;
}

static int fill_epi(const se3_l1tp_plan* p, int epilogue, int gate_ns, float cs, float cg, EpiL& e) {
    memset(&e, 0, sizeof(e));
    e.mode = epilogue;
    e.d_post = p->L.d_out;
    if (epilogue == SE3_EPI_GATE) {
        const FamL& F = p->L.f[0];
        const FamL& G = p->L.f[1];
        if (G.mz || G.mv || gate_ns < 0 || gate_ns + F.mv != F.mz) {
            set_error("gate epilogue needs outputs (ns+nv)x0e + nv x1o, got m0e=%d m1o=%d gate_ns=%d", F.mz, F.mv, gate_ns);
            return SE3_ERR_INVALID;
        }
        e.ns_g = gate_ns; e.nv = F.mv; e.d_post = gate_ns + 3 * F.mv; e.cs = cs; e.cg = cg;
    } else if (epilogue != SE3_EPI_RAW) {
        set_error("unknown epilogue %d", epilogue);
        return SE3_ERR_INVALID;
    }
    return SE3_OK;
}

static int check_weights(const se3_l1tp_plan* p, const float* const* w) {
    for (int fi = 0; fi < 2; ++fi) {
        const FamL& F = p->L.f[fi];
        if ((F.hasZ && !w[F.wz]) || (F.hasV && !w[F.wv])) { set_error("missing weight pointer"); return SE3_ERR_INVALID; }
    }
    return SE3_OK;
}

extern "C" int se3_l1tp_forward(se3_l1tp_plan* p, const se3_l1tp_fwd_args* a, void* stream) {
    if (!p || !a) { set_error("null argument"); return SE3_ERR_INVALID; }
    if (a->rows == 0) return SE3_OK;
    if (a->rows < 0 || !a->in2) { set_error("bad rows/in2"); return SE3_ERR_INVALID; }
    FwdK K;
    memset(&K, 0, sizeof(K));
    int rc = fill_src(p, a->nseg, a->seg, K.src);
    if (rc) return rc;
    rc = fill_epi(p, a->epilogue, a->gate_ns, a->gate_cs, a->gate_cg, K.epi);
    if (rc) return rc;
    rc = check_weights(p, a->w);
    if (rc) return rc;
    if (a->epilogue == SE3_EPI_GATE && a->resid) { set_error("resid is only valid with the RAW epilogue"); return SE3_ERR_INVALID; }
    if ((a->seg_idx != nullptr) != (a->out_seg != nullptr)) { set_error("seg_idx and out_seg go together"); return SE3_ERR_INVALID; }
    if (a->seg_idx && a->resid) { set_error("resid + segment sum is not supported"); return SE3_ERR_INVALID; }
    if (!a->out_raw && !a->out_post && !a->out_seg) { set_error("no output requested"); return SE3_ERR_INVALID; }
    if (a->out_post && a->epilogue != SE3_EPI_GATE) { set_error("out_post needs the GATE epilogue"); return SE3_ERR_INVALID; }
    K.rows = a->rows; K.in2 = a->in2;
    for (int s = 0; s < 4; ++s) { K.w[s] = a->w[s]; K.norm[s] = a->norm[s]; }
    K.out_raw = a->out_raw; K.out_post = a->out_post; K.resid = a->resid;
    K.seg_idx = a->seg_idx; K.out_seg = a->out_seg; K.tab = p->d_tab;
    {   // tensor-core (tcgen05) path when the configuration is eligible
        bool launched = false;
        rc = se3_l1tp_tc2_try_forward(p->n, p->m, p->t_in, p->t_out, p->L.ntab, p->h_tab.data(), p->d_tab, a, K.src, K.epi,
                                      (cudaStream_t)stream, &launched);
        if (rc) return rc;
        if (launched) return SE3_OK;
    }
    const long long ntiles = (a->rows + p->L.TR - 1) / p->L.TR;
    const int grid = (int)std::min<long long>(ntiles, (long long)num_sms() * p->occ_fwd);
    l1tp_fwd_kernel<<<grid, NT, p->L.smem_fwd, (cudaStream_t)stream>>>(p->L, K);
    SE3_LAUNCHED();
    return SE3_OK;
}

extern "C" int se3_l1tp_backward(se3_l1tp_plan* p, const se3_l1tp_bwd_args* a, void* stream) {
    if (!p || !a) { set_error("null argument"); return SE3_ERR_INVALID; }
    if (a->rows == 0) {
        for (int s = 0; s < 4; ++s)
            if (a->gw[s] && p->w_cnt[s])
                SE3_CUDA_TRY(cudaMemsetAsync(a->gw[s], 0, sizeof(float) * p->w_cnt[s], (cudaStream_t)stream));
        return SE3_OK;
    }
    if (a->rows < 0 || !a->in2 || !a->gout) { set_error("bad rows/in2/gout"); return SE3_ERR_INVALID; }
    BwdK K;
    memset(&K, 0, sizeof(K));
    int rc = fill_src(p, a->nseg, a->seg, K.src);
    if (rc) return rc;
    rc = fill_epi(p, a->epilogue, a->gate_ns, a->gate_cs, a->gate_cg, K.epi);
    if (rc) return rc;
    rc = check_weights(p, a->w);
    if (rc) return rc;
    if (a->epilogue == SE3_EPI_GATE && !a->raw) { set_error("GATE backward needs the saved pre-activation"); return SE3_ERR_INVALID; }
    if (a->epilogue == SE3_EPI_GATE && a->gin2) { set_error("gin2 is only available with the RAW epilogue"); return SE3_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    bool want_gw = false;
    for (int s = 0; s < 4; ++s) want_gw |= a->gw[s] != nullptr;
    for (int s = 0; s < a->nseg; ++s) {
        int mode = a->gseg[s] ? a->gseg_mode[s] : SE3_GRAD_NONE;
        if ((mode == SE3_GRAD_ATOMIC || mode == SE3_GRAD_SORTED) && !a->seg[s].idx) mode = SE3_GRAD_STORE;
        if (mode < SE3_GRAD_NONE || mode > SE3_GRAD_SORTED) { set_error("bad gseg_mode"); return SE3_ERR_INVALID; }
        if (mode == SE3_GRAD_STORE && a->seg[s].idx) { set_error("STORE gradient mode on an indexed segment"); return SE3_ERR_INVALID; }
        K.gseg[s] = a->gseg[s]; K.gmode[s] = mode;
    }
    K.rows = a->rows; K.in2 = a->in2;
    for (int s = 0; s < 4; ++s) { K.w[s] = a->w[s]; K.norm[s] = a->norm[s]; }
    K.raw = a->raw; K.gout = a->gout; K.gout_idx = a->gout_idx; K.tab = p->d_tab;
    const long long ntiles = (a->rows + p->L.TR - 1) / p->L.TR;
    const int grid = (int)std::min<long long>(ntiles, (long long)num_sms() * p->occ_bwd);
    int red_blocks = grid;
    bool tcw = false;
    if (want_gw) {
        const size_t need = (size_t)std::max(grid, num_sms()) * p->L.wtot;
        if (need > p->partial_cap) {
            if (p->d_partials) SE3_CUDA_TRY(cudaFree(p->d_partials));
            p->d_partials = nullptr; p->partial_cap = 0;
            const size_t cap = (size_t)num_sms() * std::max(1, p->occ_bwd) * p->L.wtot;
            SE3_CUDA_TRY(cudaMalloc(&p->d_partials, sizeof(float) * cap));
            p->partial_cap = cap;
        }
        // weight gradients on the tensor cores (accumulators resident in TMEM) when eligible
        int tc_grid = 0;
        // SE3_BWDW_GRID (diagnostics): fewer CTAs = a different partition of the rows over the fp32 accumulators
        static int grid_cap = -1;
        if (grid_cap < 0) { const char* e = getenv("SE3_BWDW_GRID"); grid_cap = e ? std::max(1, atoi(e)) : 0; }
        rc = se3_l1tp_tc2_try_backward_w(p->n, p->m, p->t_in, p->t_out, p->L.ntab, p->h_tab.data(), a, K.src, K.epi,
                                         p->d_partials, p->L.wtot, p->w_off[0], p->w_off[3],
                                         grid_cap ? std::min(grid_cap, num_sms()) : num_sms(), st, &tc_grid, &tcw);
        if (rc) return rc;
        if (tcw) red_blocks = tc_grid;
        else K.partials = p->d_partials;
    }
    bool need_in = false;
    for (int s = 0; s < a->nseg; ++s) need_in |= (K.gseg[s] != nullptr && K.gmode[s] != SE3_GRAD_NONE);
    if (need_in) {   // input gradients on the tensor cores when eligible
        bool tci = false;
        rc = se3_l1tp_tc2_try_backward_in(p->n, p->m, p->t_in, p->t_out, p->L.ntab, p->h_tab.data(), p->d_tab, a, K.src,
                                          K.epi, K.gseg, K.gmode, st, &tci);
        if (rc) return rc;
        if (tci) {
            need_in = false;
            for (int s = 0; s < SE3_MAX_SEG; ++s) { K.gseg[s] = nullptr; K.gmode[s] = SE3_GRAD_NONE; }
        }
    }
    if (need_in || K.partials) {
        if (p->maxwj == 2) l1tp_bwd_kernel<2><<<grid, NT, p->L.smem_bwd, st>>>(p->L, K);
        else l1tp_bwd_kernel<4><<<grid, NT, p->L.smem_bwd, st>>>(p->L, K);
        SE3_LAUNCHED();
    }
    if (want_gw) {
        ReduceK Rk;
        Rk.partials = p->d_partials; Rk.nblocks = red_blocks; Rk.wtot = p->L.wtot;
        for (int s = 0; s < 4; ++s) { Rk.gw[s] = a->gw[s]; Rk.off[s] = p->w_off[s]; Rk.cnt[s] = p->w_cnt[s]; }
        l1tp_reduce_gw_kernel<<<(p->L.wtot + 255) / 256, 256, 0, st>>>(Rk);
        SE3_LAUNCHED();
    }
    if (a->gin2) {
        GIn2K G;
        memset(&G, 0, sizeof(G));
        G.rows = a->rows; G.src = K.src; G.in2 = a->in2; G.gout = a->gout; G.gin2 = a->gin2; G.tab = p->d_tab;
        if (a->gout_idx) { set_error("gin2 with gout_idx is not supported"); return SE3_ERR_INVALID; }
        for (int s = 0; s < 4; ++s) { G.w[s] = a->w[s]; G.norm[s] = a->norm[s]; }
        const long long nb = (a->rows + NT / 32 - 1) / (NT / 32);
        l1tp_gin2_kernel<<<(unsigned)nb, NT, 0, st>>>(p->L, G);
        SE3_LAUNCHED();
    }
    return SE3_OK;
}
