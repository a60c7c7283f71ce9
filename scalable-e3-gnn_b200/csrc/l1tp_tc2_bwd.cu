// Second-generation tensor-core (tcgen05 / TMEM, 3xTF32) INPUT-gradient kernel of the fused l<=1 tensor-product
// layer (SEGNN case), same skeleton as l1tp_tc2.cu: 16 homogeneous worker warps + one MMA warp, operand sets and
// TMEM accumulators double buffered per 64-row tile, global -> registers -> operand tiles with no staging copy.
//
// Per row, with H the cotangent of the pre-activation (gate VJP and norms folded in):
//     T1 = [ HZ (z slots) | HG ],  HG[v] = c3 sum_c Y1[c] HV[v][c],   T3c = HV[.][c]
//     GS1 = HZ . WZ_s^T   GS2 = HG . WV_s^T   GD = HZ . WZ_d^T   GTc = HVc . WV_v^T
//     g_s[n]     = Y0 GS1[n] + GS2[n]
//     g_v[ch][c] = c3 (Y1[c] GD[ch] + Y0 GTc[ch])
// Y0 multiplies the accumulators in the epilogue instead of the operands, so HZ is stored once (the first-generation
// kernel kept both Y0*HZ and HZ tiles and had to alias its gradient tile onto them, which serialised
// build -> MMA -> epilogue -> scatter; profiles/r01_v8_bwdi).
#include <algorithm>
#include <vector>

#include "tc_common.cuh"

namespace se3 {

static constexpr int BW2 = 16;                 // worker warps
static constexpr int B2_THREADS = (BW2 + 1) * 32;
static constexpr int BWT = BW2 * 32;
static constexpr int TMB2 = 64;

struct Tc2BwdArgs {
    long long rows;
    RowSrc src;
    const float* in2;
    const float* wz;
    const float* wv;
    const float* nz;
    const float* nv;
    EpiL epi;
    const float* raw;
    const float* gout;
    const int32_t* gout_idx;
    float* gseg[SE3_MAX_SEG];
    int gmode[SE3_MAX_SEG];
    const int* tab;
    int ntab, t_s, t_d;
    int ns, nd, mz, mv, d_out, gwidth;
    int oz0, ov0;
    int nsz;                   // z channels with their own cotangent column (GATE: gate_ns, RAW: mz)
    int G0, KZ, K1T, KV;       // T1 slots: scalars 0.., gates G0.., HG KZ..;  K1T = T1 width, KV = T3 width
    int NSP, NDP;
    int nSI, nVI;
    int gts;
    int o_bs1, o_bs2, o_bd, o_bt, o_t, t_bytes, o_gt, o_norm, o_tab, o_bar, o_sidx;
    int halfT1, halfT3, oT3;
};

__device__ __forceinline__ float2 ldg2(const float* p) { return ldg2_v(p); }
__device__ __forceinline__ float4 ldg4(const float* p) { return ldg4_v(p); }

template <bool GATE>
__global__ void __launch_bounds__(B2_THREADS, 1) l1tp_tc2_bwdi_kernel(const __grid_constant__ Tc2BwdArgs A) {
    extern __shared__ __align__(1024) unsigned char smraw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int* tab = reinterpret_cast<int*>(smraw + A.o_tab);
    float* norm = reinterpret_cast<float*>(smraw + A.o_norm);     // nz[mz] then nv[3 mv]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + A.o_bar);
    const uint32_t bar0 = smem_u32(bars);
    // barriers: 0,1 operand set full | 2,3 accumulator full | 4,5 accumulator empty
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    constexpr bool gate = GATE;

    for (int t = tid; t < A.ntab; t += B2_THREADS) tab[t] = A.tab[t];
    for (int t = tid; t < A.mz; t += B2_THREADS) norm[t] = A.nz ? A.nz[t] : 1.0f;
    for (int t = tid; t < 3 * A.mv; t += B2_THREADS) norm[A.mz + t] = A.nv ? A.nv[t] : 1.0f;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(BAR(i), BW2);
            mbar_init(BAR(2 + i), 1);
            mbar_init(BAR(4 + i), BW2);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // zero both operand sets (padding slots are never written again)
        float4* z = reinterpret_cast<float4*>(smraw + A.o_t);
        const int n16 = (2 * A.t_bytes) >> 4;
        for (int t = tid; t < n16; t += B2_THREADS) z[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    {   // W^T tiles (K-major, K = T slots), hi | lo
        auto zch = [&](int k) -> int {   // T1 z slot -> z channel (or -1)
            if (k < A.nsz) return k;
            if (k >= A.G0 && k - A.G0 < A.mz - A.nsz) return A.nsz + (k - A.G0);
            return -1;
        };
        auto fill = [&](int off, int N, int K, int which) {
            unsigned char* b = smraw + off;
            const int half = N * K * 4;
            for (int t = tid; t < N * K; t += B2_THREADS) {
                const int n = t / K, k = t - n * K;
                float x = 0.0f;
                if (which == 0) { const int m = zch(k); if (n < A.ns && m >= 0) x = __ldg(A.wz + (long long)n * A.mz + m); }
                else if (which == 1) { if (n < A.ns && k < A.mv) x = __ldg(A.wv + (long long)n * A.mv + k); }
                else if (which == 2) { const int m = zch(k); if (n < A.nd && m >= 0) x = __ldg(A.wz + (long long)(A.ns + n) * A.mz + m); }
                else { if (n < A.nd && k < A.mv) x = __ldg(A.wv + (long long)(A.ns + n) * A.mv + k); }
                float hi, lo;
                split_tf32(x, hi, lo);
                const int o = canon_off(n, k, K >> 2);
                *reinterpret_cast<float*>(b + o) = hi;
                *reinterpret_cast<float*>(b + half + o) = lo;
            }
        };
        fill(A.o_bs1, A.NSP, A.KZ, 0);
        fill(A.o_bs2, A.NSP, A.K1T - A.KZ, 1);
        fill(A.o_bd, A.NDP, A.KZ, 2);
        fill(A.o_bt, A.NDP, A.KV, 3);
    }
    fence_proxy_async();
    if (warp == BW2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const long long R = A.rows;
    const long long ntiles = (R + TMB2 - 1) / TMB2;
    const int nt = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const uint32_t ACC = 256;
    const int cS1 = 0, cS2 = A.NSP, cD = 2 * A.NSP, cT = 2 * A.NSP + A.NDP;

    if (warp == BW2) {
        // ================= MMA issuer
        const uint32_t sb = smem_u32(smraw);
        const uint32_t idS = make_idesc(A.NSP), idD = make_idesc(A.NDP);
        const int K2 = A.K1T - A.KZ;
        const uint32_t sboT1 = (A.K1T >> 2) * 128, sboT3 = (A.KV >> 2) * 128;
        const uint32_t sboZ = (A.KZ >> 2) * 128, sboG = (K2 >> 2) * 128;
        const uint64_t bS1h = make_desc(sb + A.o_bs1, sboZ), bS1l = make_desc(sb + A.o_bs1 + A.NSP * A.KZ * 4, sboZ);
        const uint64_t bS2h = make_desc(sb + A.o_bs2, sboG), bS2l = make_desc(sb + A.o_bs2 + A.NSP * K2 * 4, sboG);
        const uint64_t bDh = make_desc(sb + A.o_bd, sboZ), bDl = make_desc(sb + A.o_bd + A.NDP * A.KZ * 4, sboZ);
        const uint64_t bTh = make_desc(sb + A.o_bt, sboT3), bTl = make_desc(sb + A.o_bt + A.NDP * A.KV * 4, sboT3);
        const int nkz = A.KZ >> 3, nkg = K2 >> 3, nkv = A.KV >> 3;
        const uint64_t v3 = (uint64_t)((2u * A.halfT3) >> 4);
        for (int it = 0; it < nt; ++it) {
            const int b = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            mbar_wait(BAR(b), ph);
            mbar_wait(BAR(4 + b), ph ^ 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t acc = tmem_base + (uint32_t)b * ACC;
                const uint32_t t1 = sb + A.o_t + (uint32_t)b * A.t_bytes;
                const uint64_t a1h = make_desc(t1, sboT1), a1l = make_desc(t1 + A.halfT1, sboT1);
                for (int j = 0; j < nkz; ++j) {
                    const uint64_t o = (uint64_t)(j * 16);
                    tc_mma_tf32(acc + cS1, a1h + o, bS1h + o, idS, j ? 1u : 0u);
                    tc_mma_tf32(acc + cS1, a1h + o, bS1l + o, idS, 1u);
                    tc_mma_tf32(acc + cS1, a1l + o, bS1h + o, idS, 1u);
                    tc_mma_tf32(acc + cD, a1h + o, bDh + o, idD, j ? 1u : 0u);
                    tc_mma_tf32(acc + cD, a1h + o, bDl + o, idD, 1u);
                    tc_mma_tf32(acc + cD, a1l + o, bDh + o, idD, 1u);
                }
                for (int j = 0; j < nkg; ++j) {
                    const uint64_t o = (uint64_t)(j * 16), oa = (uint64_t)((nkz + j) * 16);
                    tc_mma_tf32(acc + cS2, a1h + oa, bS2h + o, idS, j ? 1u : 0u);
                    tc_mma_tf32(acc + cS2, a1h + oa, bS2l + o, idS, 1u);
                    tc_mma_tf32(acc + cS2, a1l + oa, bS2h + o, idS, 1u);
                }
                const uint64_t a3h = make_desc(t1 + A.oT3, sboT3), a3l = make_desc(t1 + A.oT3 + A.halfT3, sboT3);
                for (int c = 0; c < 3; ++c)
                    for (int j = 0; j < nkv; ++j) {
                        const uint64_t o = (uint64_t)(j * 16), oa = o + (uint64_t)c * v3;
                        tc_mma_tf32(acc + cT + c * A.NDP, a3h + oa, bTh + o, idD, j ? 1u : 0u);
                        tc_mma_tf32(acc + cT + c * A.NDP, a3h + oa, bTl + o, idD, 1u);
                        tc_mma_tf32(acc + cT + c * A.NDP, a3l + oa, bTh + o, idD, 1u);
                    }
                tc_commit(BAR(2 + b));
            }
            __syncwarp();
        }
    } else {
        // ================= workers.  warp w: row group rb = w & 7 (rows 8 rb + (lane & 7)), item block jb = w >> 3.
        //   S item  j = 4 jb + jq           : z channels 4j..4j+3 without a gate role (swish VJP / plain)
        //   X item  q = 4 jb + jq           : q < nVI -> vector channels 2q, 2q+1 (+ their gate scalars);
        //                                     nVI <= q < nVI + (nSI - 8) -> S item 8 + q - nVI
        const int r8 = lane & 7, jq = lane >> 3;
        const int wrow = (warp & 7) * 8 + r8;
        const int jb = warp >> 3;
        const int KQ1 = A.K1T >> 2, KQ3 = A.KV >> 2;
        const int rp1 = (((wrow >> 3) * KQ1) << 7) + ((wrow & 7) << 4);
        const int rp3 = (((wrow >> 3) * KQ3) << 7) + ((wrow & 7) << 4);
        const int sj = 4 * jb + jq;
        const bool s_act = sj < A.nSI && sj < 8;
        const int xq = 4 * jb + jq;
        const bool v_act = xq < A.nVI;
        const int sj2 = 8 + xq - A.nVI;
        const bool s2_act = !v_act && sj2 >= 8 && sj2 < A.nSI;
        const float cs = A.epi.cs, cg = A.epi.cg;
        // prefetched registers of the next tile
        float4 sR, sG, s2R, s2G;          // S items: raw / cotangent of 4 channels
        float2 vRg, vRv[3], vGv[3];       // V item: raw gate pair, raw vectors (6), cotangent vectors (6)
        float4 yv;
        long long growi = 0, gidx = 0, growi_n = 0, gidx_n = 0;
        auto load_idx = [&](int it) {   // row / cotangent-row index of tile `it`, fetched one tile ahead of the row loads
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * TMB2;
            long long gr = row0 + wrow;
            if (gr > R - 1) gr = R - 1;
            growi_n = gr;
            gidx_n = A.gout_idx ? (long long)ldgi_v(A.gout_idx + gr) : gr;
        };
        auto load_s = [&](int j, float4& r4, float4& g4) {
            const int c = 4 * j;
            if (gate) {
                const float* rp = A.raw + growi * A.d_out + A.oz0 + c;
                const float2 a = ldg2(rp), b = ldg2(rp + 2);
                r4 = make_float4(a.x, a.y, b.x, b.y);
                g4 = ldg4(A.gout + gidx * A.gwidth + c);
            } else {
                const float* gp = A.gout + gidx * A.gwidth + A.oz0 + c;
                const float2 a = ldg2(gp), b = ldg2(gp + 2);
                g4 = make_float4(a.x, a.y, b.x, b.y);
                r4 = g4;
            }
        };
        auto load_rows = [&]() {
            if (s_act) load_s(sj, sR, sG);
            if (s2_act) load_s(sj2, s2R, s2G);
            if (v_act) {
                if (gate) {
                    vRg = ldg2(A.raw + growi * A.d_out + A.oz0 + A.nsz + 2 * xq);
                    const float* rv = A.raw + growi * A.d_out + A.ov0 + 6 * xq;
                    const float* gv = A.gout + gidx * A.gwidth + A.nsz + 6 * xq;
#pragma unroll
                    for (int u = 0; u < 3; ++u) { vRv[u] = ldg2(rv + 2 * u); vGv[u] = ldg2(gv + 2 * u); }
                } else {
                    const float* gv = A.gout + gidx * A.gwidth + A.ov0 + 6 * xq;
#pragma unroll
                    for (int u = 0; u < 3; ++u) vGv[u] = ldg2(gv + 2 * u);
                }
                yv = ldg4(A.in2 + growi * 4);
            }
        };
        auto put4 = [&](unsigned char* base, int half, int off, float a, float b, float c, float d) {
            float4 h, l;
            split_tf32(a, h.x, l.x); split_tf32(b, h.y, l.y); split_tf32(c, h.z, l.z); split_tf32(d, h.w, l.w);
            *reinterpret_cast<float4*>(base + off) = h;
            *reinterpret_cast<float4*>(base + half + off) = l;
        };
        auto put2 = [&](unsigned char* base, int half, int off, float a, float b) {
            float2 h, l;
            split_tf32(a, h.x, l.x); split_tf32(b, h.y, l.y);
            *reinterpret_cast<float2*>(base + off) = h;
            *reinterpret_cast<float2*>(base + half + off) = l;
        };
        const float* nzs = norm;
        const float* nvs = norm + A.mz;
        auto build_s = [&](unsigned char* tset, int j, const float4& r4, const float4& g4) {
            const int m0 = 4 * j;
            float h[4];
            const float rr[4] = {r4.x, r4.y, r4.z, r4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float v = 0.0f;
                if (m0 + u < A.nsz) {
                    if (gate) {
                        const float x = rr[u], s = sigm(x);
                        v = gg[u] * cs * s * (1.0f + x * (1.0f - s)) * nzs[m0 + u];
                    } else {
                        v = gg[u] * nzs[m0 + u];
                    }
                }
                h[u] = v;
            }
            put4(tset, A.halfT1, rp1 + (j << 7), h[0], h[1], h[2], h[3]);
        };
        auto build = [&](int b) {
            unsigned char* tset = smraw + A.o_t + b * A.t_bytes;
            if (s_act) build_s(tset, sj, sR, sG);
            if (s2_act) build_s(tset, sj2, s2R, s2G);
            if (v_act) {
                const int v0 = 2 * xq;
                const float gv[6] = {vGv[0].x, vGv[0].y, vGv[1].x, vGv[1].y, vGv[2].x, vGv[2].y};
                float hv[6], hz[2] = {0.f, 0.f}, hg[2];
                if (gate) {
                    const float rv[6] = {vRv[0].x, vRv[0].y, vRv[1].x, vRv[1].y, vRv[2].x, vRv[2].y};
                    const float rg[2] = {vRg.x, vRg.y};
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const float s = sigm(rg[u]);
                        const float dot = gv[3 * u] * rv[3 * u] + gv[3 * u + 1] * rv[3 * u + 1] + gv[3 * u + 2] * rv[3 * u + 2];
                        hz[u] = cg * s * (1.0f - s) * dot * nzs[A.nsz + v0 + u];
                        const float sg = cg * s;
#pragma unroll
                        for (int c = 0; c < 3; ++c) hv[3 * u + c] = sg * gv[3 * u + c] * nvs[3 * (v0 + u) + c];
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < 2; ++u)
#pragma unroll
                        for (int c = 0; c < 3; ++c) hv[3 * u + c] = gv[3 * u + c] * nvs[3 * (v0 + u) + c];
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (v0 + u >= A.mv) { hz[u] = 0.f; hv[3 * u] = hv[3 * u + 1] = hv[3 * u + 2] = 0.f; }
                    hg[u] = C3f * (yv.y * hv[3 * u] + yv.z * hv[3 * u + 1] + yv.w * hv[3 * u + 2]);
                }
                // slots: gate scalars G0 + v, HG KZ + v (both pairs are 8-byte aligned: G0, KZ, v0 even), HV_c v
                if (gate) put2(tset, A.halfT1, rp1 + (((A.G0 + v0) >> 2) << 7) + (((A.G0 + v0) & 3) << 2), hz[0], hz[1]);
                put2(tset, A.halfT1, rp1 + (((A.KZ + v0) >> 2) << 7) + (((A.KZ + v0) & 3) << 2), hg[0], hg[1]);
                const int o3 = A.oT3 + rp3 + ((v0 >> 2) << 7) + ((v0 & 3) << 2);
#pragma unroll
                for (int c = 0; c < 3; ++c) put2(tset, A.halfT3, o3 + c * 2 * A.halfT3, hv[c], hv[3 + c]);
            }
        };
        // epilogue part 1: TMEM -> gradient tile (flat concatenated in1 layout) in shared memory
        const int e = warp & 3, cgq = warp >> 2;
        float* gt = reinterpret_cast<float*>(smraw + A.o_gt);
        const int gts = A.gts;
        const int drow = 16 * e + (lane & 15);
        const bool rowlane = lane < 16;
        const int* scol = tab + A.t_s;
        const int* vcol = tab + A.t_d;
        int* sidx = reinterpret_cast<int*>(smraw + A.o_sidx);   // [SE3_MAX_SEG][64] destination rows of the tile being scattered
        int pidx[SE3_MAX_SEG];
        float4 ypre = make_float4(0.f, 0.f, 0.f, 0.f), ypre2 = ypre;   // in2 of this thread's two epilogue rows
        auto prefetch_sidx = [&](int it) {   // issued early in the iteration; consumed by drain(it)
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * TMB2;
            long long gr = row0 + drow;
            if (gr > R - 1) gr = R - 1;
            {
                long long ga = row0 + 16 * e + (lane >> 2), gb = ga + 8;
                if (ga > R - 1) ga = R - 1;
                if (gb > R - 1) gb = R - 1;
                ypre = ldg4_v(A.in2 + ga * 4);
                ypre2 = ldg4_v(A.in2 + gb * 4);
            }
#pragma unroll
            for (int s = 0; s < SE3_MAX_SEG; ++s) {
                pidx[s] = (int)gr;
                if (cgq == 0 && rowlane && s < A.src.nseg && A.src.idx[s]) pidx[s] = ldgi_v(A.src.idx[s] + gr);
            }
        };
        const int fg = lane >> 2, fq = lane & 3;   // 16x256b fragment: rows fg, fg + 8 of the quarter; columns 2 fq, 2 fq + 1
        auto drain = [&](int it) {
            const int b = it & 1;
            const float4 y = ypre, y2 = ypre2;
            const float ya[4] = {C3f * y.x, C3f * y.y, C3f * y.z, C3f * y.w};
            const float yb[4] = {C3f * y2.x, C3f * y2.y, C3f * y2.z, C3f * y2.w};
            mbar_wait(BAR(2 + b), (it >> 1) & 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + (uint32_t)b * ACC + ((uint32_t)(32 * e) << 16);
            float* rowa = gt + (16 * e + fg) * gts;
            float* rowb = rowa + 8 * gts;
            for (int k0 = 8 * cgq; k0 < A.NSP; k0 += 8 * (BW2 / 4)) {
                float a[4], c[4];
                tc_ld_16x256(acc + cS1 + k0, a);
                tc_ld_16x256(acc + cS2 + k0, c);
                tc_wait_ld();
                const int n = k0 + 2 * fq;
#pragma unroll
                for (int u = 0; u < 2; ++u)
                    if (n + u < A.ns) {
                        const int col = scol[n + u];
                        rowa[col] = fmaf(y.x, a[u], c[u]);
                        rowb[col] = fmaf(y2.x, a[2 + u], c[2 + u]);
                    }
            }
            for (int k0 = 8 * cgq; k0 < A.NDP; k0 += 8 * (BW2 / 4)) {
                float d[4], t0[4], t1[4], t2[4];
                tc_ld_16x256(acc + cD + k0, d);
                tc_ld_16x256(acc + cT + k0, t0);
                tc_ld_16x256(acc + cT + A.NDP + k0, t1);
                tc_ld_16x256(acc + cT + 2 * A.NDP + k0, t2);
                tc_wait_ld();
                const int ch = k0 + 2 * fq;
#pragma unroll
                for (int u = 0; u < 2; ++u)
                    if (ch + u < A.nd) {
                        const int col = vcol[ch + u];
                        float* oa = rowa + col;
                        float* ob = rowb + col;
                        oa[0] = fmaf(ya[1], d[u], ya[0] * t0[u]);
                        oa[1] = fmaf(ya[2], d[u], ya[0] * t1[u]);
                        oa[2] = fmaf(ya[3], d[u], ya[0] * t2[u]);
                        ob[0] = fmaf(yb[1], d[2 + u], yb[0] * t0[2 + u]);
                        ob[1] = fmaf(yb[2], d[2 + u], yb[0] * t1[2 + u]);
                        ob[2] = fmaf(yb[3], d[2 + u], yb[0] * t2[2 + u]);
                    }
            }
            if (cgq == 0 && rowlane) {
#pragma unroll
                for (int s = 0; s < SE3_MAX_SEG; ++s) sidx[s * TMB2 + drow] = pidx[s];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(4 + b));
        };
        // epilogue part 2: scatter the gradient tile per segment
        auto scatter = [&](int it) {
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * TMB2;
            const int nvalid = (int)min((long long)TMB2, R - row0);
#pragma unroll
            for (int s = 0; s < SE3_MAX_SEG; ++s) {   // unrolled: no dynamic indexing of the kernel parameters in the tile loop
                if (s >= A.src.nseg) break;
                float* gb = A.gseg[s];
                const int mode = A.gmode[s];
                if (!gb || mode == SE3_GRAD_NONE) continue;
                const int w = A.src.cum[s + 1] - A.src.cum[s], c0 = A.src.cum[s], ld = A.src.ld[s];
                const int* idx = sidx + s * TMB2;   // destination row of every tile row (identity rows included)
                const bool v4 = (w & 3) == 0 && (ld & 3) == 0 && (c0 & 3) == 0 && ((uintptr_t)gb & 15) == 0;
                if (mode == SE3_GRAD_STORE || mode == SE3_GRAD_ATOMIC) {
                    if (v4) {
                        const int w4 = w >> 2;
                        for (int t = tid; t < nvalid * w4; t += BWT) {
                            const int r = t / w4, c = (t - r * w4) << 2;
                            const float4 v = *reinterpret_cast<const float4*>(gt + r * gts + c0 + c);
                            const long long dr = idx[r];
                            float* dst = gb + dr * ld + c;
                            if (mode == SE3_GRAD_STORE) *reinterpret_cast<float4*>(dst) = v;
                            else red_add_v4(dst, v.x, v.y, v.z, v.w);
                        }
                    } else {
                        for (int t = tid; t < nvalid * w; t += BWT) {
                            const int r = t / w, c = t - r * w;
                            const float v = gt[r * gts + c0 + c];
                            const long long dr = idx[r];
                            if (mode == SE3_GRAD_STORE) gb[dr * ld + c] = v;
                            else atomicAdd(gb + dr * ld + c, v);
                        }
                    }
                } else {  // SORTED: run-length combine equal destinations, one red per run
                    int parts = BWT / w;
                    if (parts < 1) parts = 1;
                    if (parts > TMB2) parts = TMB2;
                    const int rpp = (TMB2 + parts - 1) / parts;
                    for (int item = tid; item < w * parts; item += BWT) {
                        const int c = item % w, qd = item / w;
                        const int rbeg = qd * rpp;
                        const int rend = min(rbeg + rpp, nvalid);
                        if (rbeg >= rend) continue;
                        int cur = idx[rbeg];
                        float accv = 0.0f;
                        for (int r = rbeg; r < rend; ++r) {
                            const int k = idx[r];
                            if (k != cur) {
                                atomicAdd(gb + (long long)cur * ld + c, accv);
                                cur = k;
                                accv = 0.0f;
                            }
                            accv += gt[r * gts + c0 + c];
                        }
                        atomicAdd(gb + (long long)cur * ld + c, accv);
                    }
                }
            }
        };

        if (nt > 0) {
            load_idx(0);
            growi = growi_n; gidx = gidx_n;
            load_rows();
            if (nt > 1) load_idx(1);
        }
        for (int it = 0; it < nt; ++it) {
            build(it & 1);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(it & 1));
            if (it + 1 < nt) {
                growi = growi_n; gidx = gidx_n;
                load_rows();
                if (it + 2 < nt) load_idx(it + 2);
            }
            if (it >= 1) {
                named_bar(2, BWT);
                drain(it - 1);
                named_bar(1, BWT);
                scatter(it - 1);
            }
            prefetch_sidx(it);
        }
        if (nt > 0) {
            named_bar(2, BWT);
            drain(nt - 1);
            named_bar(1, BWT);
            scatter(nt - 1);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == BW2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

}  // namespace se3

using namespace se3;

int se3_l1tp_tc2_try_backward_in(const int n[4], const int m[4], const int t_in[4], const int t_out[4], int ntab,
                                 const int* h_tab, const int* d_tab, const se3_l1tp_bwd_args* a, const RowSrc& src,
                                 const EpiL& epi, float* const gseg[SE3_MAX_SEG], const int gmode[SE3_MAX_SEG],
                                 cudaStream_t st, bool* launched) {
    *launched = false;
    static int disabled = -1;
    if (disabled < 0) {
        const char* e = getenv("SE3_DISABLE_TC2");
        const char* e1 = getenv("SE3_DISABLE_TC");
        disabled = ((e && (e[0] == '1' || e[0] == '3')) || (e1 && (e1[0] == '1' || e1[0] == '3'))) ? 1 : 0;
    }
    if (disabled) return SE3_OK;
    if (n[1] || n[2] || m[1] || m[2]) return SE3_OK;
    const int ns = n[0], nd = n[3], mz = m[0], mv = m[3];
    if (ns < 1 || nd < 1 || mz < 1 || mv < 1) return SE3_OK;
    if (a->rows >= (1ll << 31) - TMB2) return SE3_OK;
    static thread_local Tc2BwdArgs A;
    memset(&A, 0, sizeof(A));
    const bool gate = epi.mode == SE3_EPI_GATE;
    A.ns = ns; A.nd = nd; A.mz = mz; A.mv = mv; A.d_out = mz + 3 * mv;
    A.oz0 = h_tab[t_out[0]]; A.ov0 = h_tab[t_out[3]];
    for (int k = 0; k < mz; ++k) if (h_tab[t_out[0] + k] != A.oz0 + k) return SE3_OK;
    for (int k = 0; k < mv; ++k) if (h_tab[t_out[3] + k] != A.ov0 + 3 * k) return SE3_OK;
    A.nsz = gate ? epi.ns_g : mz;
    if (gate && (epi.ns_g < 1 || epi.ns_g + mv != mz)) return SE3_OK;
    A.gwidth = epi.d_post;
    // alignment of the 8- and 16-byte loads (see the kernel): everything even, cotangent rows 16-byte aligned when gated
    if ((mv & 1) || (A.nsz & 1) || (A.oz0 & 1) || (A.ov0 & 1) || (A.d_out & 1) || (A.gwidth & 1)) return SE3_OK;
    if (((uintptr_t)a->in2 & 15) || ((uintptr_t)a->gout & 15) || (gate && ((uintptr_t)a->raw & 7))) return SE3_OK;
    if (gate && (A.gwidth & 3)) return SE3_OK;
    A.nSI = (A.nsz + 3) >> 2;
    A.nVI = mv >> 1;
    if (A.nVI > 8 || A.nSI > 8 + (8 - A.nVI)) return SE3_OK;
    A.G0 = (A.nsz + 3) & ~3;
    A.KZ = gate ? ((A.G0 + mv + 7) & ~7) : ((mz + 7) & ~7);
    A.K1T = (A.KZ + mv + 7) & ~7;
    A.KV = (mv + 7) & ~7;
    A.NSP = (ns + 7) & ~7; A.NDP = (nd + 7) & ~7;
    if (2 * A.NSP + 4 * A.NDP > 256 || A.NSP > 256) return SE3_OK;
    A.rows = a->rows; A.src = src; A.in2 = a->in2; A.wz = a->w[0]; A.wv = a->w[3]; A.nz = a->norm[0]; A.nv = a->norm[3];
    A.epi = epi; A.raw = a->raw; A.gout = a->gout; A.gout_idx = a->gout_idx; A.tab = d_tab; A.ntab = ntab;
    A.t_s = t_in[0]; A.t_d = t_in[3];
    for (int s = 0; s < SE3_MAX_SEG; ++s) { A.gseg[s] = gseg[s]; A.gmode[s] = gmode[s]; }
    auto al = [](int x, int q) { return (x + q - 1) / q * q; };
    A.halfT1 = TMB2 * A.K1T * 4; A.halfT3 = TMB2 * A.KV * 4; A.oT3 = 2 * A.halfT1;
    A.t_bytes = 2 * A.halfT1 + 6 * A.halfT3;
    A.gts = (src.cum[src.nseg] + 3) & ~3;
    if ((A.gts & 31) == 0) A.gts += 4;
    int o = 0;
    A.o_bs1 = o; o += 2 * A.NSP * A.KZ * 4;
    A.o_bs2 = o; o += 2 * A.NSP * (A.K1T - A.KZ) * 4;
    A.o_bd = o; o += 2 * A.NDP * A.KZ * 4;
    A.o_bt = o; o += 2 * A.NDP * A.KV * 4;
    o = al(o, 1024);
    A.o_t = o; o += 2 * A.t_bytes;
    A.o_gt = o; o += al(TMB2 * A.gts * 4, 16);
    A.o_norm = o; o += al((mz + 3 * mv) * 4, 16);
    A.o_tab = o; o += al(ntab * 4, 16);
    A.o_bar = o; o += 8 * 8 + 16;
    A.o_sidx = o; o += SE3_MAX_SEG * TMB2 * 4;
    int dev = 0, maxsm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (o > maxsm) return SE3_OK;
    const int smem = std::max(o, 120 * 1024);
    static bool attr_set = false;
    if (!attr_set) {
        SE3_CUDA_TRY(cudaFuncSetAttribute(l1tp_tc2_bwdi_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
        SE3_CUDA_TRY(cudaFuncSetAttribute(l1tp_tc2_bwdi_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
        attr_set = true;
    }
    const long long ntiles = (a->rows + TMB2 - 1) / TMB2;
    const int grid = (int)std::min<long long>(ntiles, num_sms());
    if (gate) l1tp_tc2_bwdi_kernel<true><<<grid, B2_THREADS, smem, st>>>(A);
    else l1tp_tc2_bwdi_kernel<false><<<grid, B2_THREADS, smem, st>>>(A);
    SE3_LAUNCHED();
    g_tc_launches.fetch_add(1, std::memory_order_relaxed);
    *launched = true;
    return SE3_OK;
}
