// Second-generation tensor-core (tcgen05 / TMEM, 3xTF32) forward of the fused l<=1 tensor-product layer, SEGNN
// case (inputs a x0e + b x1o in up to four gathered segments, outputs c x0e + d x1o).
//
// What changed against l1tp_tc.cu (profiles/r01_v8_*: the MMA-issuing warp and the per-chunk hand-offs were the
// bottleneck, 11.5k cycles per 64-row tile against a tensor-pipe floor of 1.3k):
//   * no staging copy and no role pipeline: 16 homogeneous worker warps gather the rows of tile t+1 from global
//     memory straight into registers (issued a whole tile ahead), split them into tf32 hi/lo and store them ONCE, in
//     their final UMMA operand form; one barrier per TILE hands the operand set to the MMA warp;
//   * the operand set holds only raw data: scalars S[64 x K1] and the de-interleaved vector components
//     Vx, Vy, Vz [64 x K2].  Nothing in it depends on the spherical harmonics, which enter in the epilogue:
//         P  = S  . [WZ_s | WV_s]           (N1 columns)
//         Uc = Vc . [WZ_d | WV_v]  c=x,y,z  (3 x N1 columns)
//         out0[m]    = norm (Y0 P[m] + c3 sum_c Y1[c] Uc[m])
//         out1[m][c] = norm c3 (Y1[c] P[N2+m] + Y0 Uc[N2+m])
//     so the dot-product / Y0-scaled feature tiles of the first kernel (and their builders) are gone;
//   * operand sets and TMEM accumulators are double buffered per tile: the MMAs of tile t run while the workers
//     build tile t+1 and finish tile t-1 (TMEM -> registers -> smem tile -> coalesced stores / gate / segment sum);
//   * all ring arithmetic is 32-bit and additive; the MMA warp only adds constants to precomputed descriptors.
#include <algorithm>
#include <vector>

#include "tc_common.cuh"

namespace se3 {

static constexpr int W2 = 16;                 // worker warps (multiple of 8: row groups; of 4: TMEM lane quarters)
static constexpr int T2_THREADS = (W2 + 1) * 32;
static constexpr int WT = W2 * 32;            // worker threads
static constexpr int TM2 = 64;
static constexpr int MAXCOL = 192;
static constexpr int MAXK1 = 128;

static constexpr int NG = W2 / 8;              // warps per row group
static constexpr int MAXB = 4;                // wide blocks per warp and tile (NG * MAXB per tile)

struct BlockE {                               // one wide block: 64 rows x 4 consecutive 16-byte pieces of a segment
    const float* base;
    const int32_t* idx;
    int ld;
    int cum;                                  // first concatenated column of the segment
    int nch;                                  // 16-byte pieces per row of the segment
    int cb;                                   // pieces 4 cb .. 4 cb + 3
    unsigned fast4;                           // byte q = K-chunk + 1 of piece 4 cb + q if it is 4 aligned scalars
    int pad;
};
struct NarrowE {                              // 32 four-byte pieces of a narrow segment (row-major over 64 rows x w)
    const float* base;
    const int32_t* idx;
    int ld, cum, w, t;
};

struct Tc2Args {
    long long rows;
    const float* in2;
    const float* wz;
    const float* wv;
    const float* nz;
    const float* nv;
    EpiL epi;
    float* out_raw;
    float* out_post;
    const float* resid;
    const int32_t* seg_idx;
    float* out_seg;
    int ns, nd, mz, mv, d_out;
    int oz0, ov0;                             // affine output layout: z channel m at oz0 + m, v channel m at ov0 + 3 m
    int K1, K2, N1, N2;
    int nblk, nnar;
    int dop, dpp;
    unsigned mg_ns, mg_mv, mg_h;              // ceil(2^32 / d) for d = gate_ns, mv, d_out / 2
    int o_b1, o_b2, o_a, a_bytes, o_out, o_post, o_norm, o_blk, o_ccode, o_bar, o_sseg, o_hs;
    int halfS, halfV, oV, HALF;               // bytes of one S / V tile (hi), first V tile, lo offset of every tile (= all hi tiles)
    BlockE blk[NG * MAXB];
    NarrowE nar[W2];
    unsigned short ccode[MAXCOL];             // per concatenated column: type << 13 | index (1: scalar slot, 2..4: x/y/z kd)
    short sl2ch[MAXK1];                       // scalar slot -> scalar channel (-1: padding)
};

// SEG: the launch ends in the sorted-segment sum (kept out of the other instantiation: its prefetch registers and
// barriers cost the plain launches 7 % through register pressure)
template <bool SEG, bool GATE>
__global__ void __launch_bounds__(T2_THREADS, 1) l1tp_tc2_fwd_kernel(const __grid_constant__ Tc2Args A) {
    extern __shared__ __align__(1024) unsigned char smraw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* norm = reinterpret_cast<float*>(smraw + A.o_norm);
    BlockE* blks = reinterpret_cast<BlockE*>(smraw + A.o_blk);
    unsigned short* ccode = reinterpret_cast<unsigned short*>(smraw + A.o_ccode);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + A.o_bar);
    const uint32_t bar0 = smem_u32(bars);
    // barriers: 0,1 operand set full | 2,3 accumulator full | 4,5 accumulator empty
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    // ---------------- one-time setup
    for (int t = tid; t < 3 * A.mv; t += T2_THREADS) norm[t] = A.nv ? A.nv[t] : 1.0f;
    for (int t = tid; t < A.nblk; t += T2_THREADS) blks[t] = A.blk[t];
    for (int t = tid; t < MAXCOL; t += T2_THREADS) ccode[t] = A.ccode[t];
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(BAR(i), W2);
            mbar_init(BAR(2 + i), 1);
            mbar_init(BAR(4 + i), W2);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // zero both operand sets (padding slots are never written again)
        float4* z = reinterpret_cast<float4*>(smraw + A.o_a);
        const int n16 = (2 * A.a_bytes) >> 4;
        for (int t = tid; t < n16; t += T2_THREADS) z[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    {   // weights -> canonical K-major B tiles (hi | lo)
        const int KQ1 = A.K1 >> 2, KQ2 = A.K2 >> 2;
        unsigned char* b1 = smraw + A.o_b1;
        const int half1 = A.N1 * A.K1 * 4;
        for (int t = tid; t < A.N1 * A.K1; t += T2_THREADS) {
            const int n = t / A.K1, k = t - n * A.K1;
            const int ch = A.sl2ch[k];
            float x = 0.0f;
            if (ch >= 0) {
                if (n < A.mz) x = __ldg(A.wz + (long long)ch * A.mz + n) * (A.nz ? __ldg(A.nz + n) : 1.0f);
                else if (n >= A.N2 && n - A.N2 < A.mv) x = __ldg(A.wv + (long long)ch * A.mv + (n - A.N2));
            }
            float hi, lo;
            split_tf32(x, hi, lo);
            const int o = canon_off(n, k, KQ1);
            *reinterpret_cast<float*>(b1 + o) = hi;
            *reinterpret_cast<float*>(b1 + half1 + o) = lo;
        }
        unsigned char* b2 = smraw + A.o_b2;
        const int half2 = A.N1 * A.K2 * 4;
        for (int t = tid; t < A.N1 * A.K2; t += T2_THREADS) {
            const int n = t / A.K2, k = t - n * A.K2;
            float x = 0.0f;
            if (k < A.nd) {
                if (n < A.mz) x = __ldg(A.wz + (long long)(A.ns + k) * A.mz + n) * (A.nz ? __ldg(A.nz + n) : 1.0f);
                else if (n >= A.N2 && n - A.N2 < A.mv) x = __ldg(A.wv + (long long)(A.ns + k) * A.mv + (n - A.N2));
            }
            float hi, lo;
            split_tf32(x, hi, lo);
            const int o = canon_off(n, k, KQ2);
            *reinterpret_cast<float*>(b2 + o) = hi;
            *reinterpret_cast<float*>(b2 + half2 + o) = lo;
        }
    }
    fence_proxy_async();
    if (warp == W2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const long long R = A.rows;
    const long long ntiles = (R + TM2 - 1) / TM2;
    const int nt = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);   // tiles of this CTA
    const uint32_t ACC = 4u * A.N1;                                            // TMEM columns per accumulator set

    if (warp == W2) {
        // ================= MMA issuer
        const uint32_t sb = smem_u32(smraw);
        const uint32_t idesc = make_idesc(A.N1);
        const uint32_t sboS = (A.K1 >> 2) * 128, sboV = (A.K2 >> 2) * 128;
        const uint64_t dB1h = make_desc(sb + A.o_b1, sboS), dB1l = make_desc(sb + A.o_b1 + A.N1 * A.K1 * 4, sboS);
        const uint64_t dB2h = make_desc(sb + A.o_b2, sboV), dB2l = make_desc(sb + A.o_b2 + A.N1 * A.K2 * 4, sboV);
        const int nk1 = A.K1 >> 3, nk2 = A.K2 >> 3;
        const uint32_t vstep = ((uint32_t)A.halfV) >> 4;   // descriptor units between consecutive V tiles
        for (int it = 0; it < nt; ++it) {
            const int b = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            mbar_wait(BAR(b), ph);
            mbar_wait(BAR(4 + b), ph ^ 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t acc = tmem_base + (uint32_t)b * ACC;
                const uint32_t aS = sb + A.o_a + (uint32_t)b * A.a_bytes;
                const uint64_t dSh = make_desc(aS, sboS), dSl = make_desc(aS + A.HALF, sboS);
                for (int j = 0; j < nk1; ++j) {
                    const uint64_t o = (uint64_t)(j * 16);   // 256 bytes per K-step, in 16-byte units
                    tc_mma_tf32(acc, dSh + o, dB1h + o, idesc, j ? 1u : 0u);
                    tc_mma_tf32(acc, dSh + o, dB1l + o, idesc, 1u);
                    tc_mma_tf32(acc, dSl + o, dB1h + o, idesc, 1u);
                }
                const uint64_t dVh = make_desc(aS + A.oV, sboV), dVl = make_desc(aS + A.oV + A.HALF, sboV);
                for (int c = 0; c < 3; ++c) {
                    const uint32_t accc = acc + (uint32_t)(c + 1) * A.N1;
                    const uint64_t co = (uint64_t)c * vstep;
                    for (int j = 0; j < nk2; ++j) {
                        const uint64_t o = (uint64_t)(j * 16);
                        tc_mma_tf32(accc, dVh + co + o, dB2h + o, idesc, j ? 1u : 0u);
                        tc_mma_tf32(accc, dVh + co + o, dB2l + o, idesc, 1u);
                        tc_mma_tf32(accc, dVl + co + o, dB2h + o, idesc, 1u);
                    }
                }
                tc_commit(BAR(2 + b));
            }
            __syncwarp();
        }
    } else {
        // ================= workers
        // Everything about WHERE a thread's pieces come from and go to is tile-invariant and lives in registers:
        // warp w owns row group rb = w & 7 (rows 8 rb + (lane & 7)) of the wide blocks (w >> 3) + NG i.
        const int KQ1 = A.K1 >> 2, KQ2 = A.K2 >> 2;
        const int r8 = lane & 7, cq = lane >> 3;
        const int wrow = (warp & 7) * 8 + r8;
        const int rowpartS = (((wrow >> 3) * KQ1) << 7) + ((wrow & 7) << 4);
        const int rowpartV = (((wrow >> 3) * KQ2) << 7) + ((wrow & 7) << 4);
        auto enc = [&](int code, int rpS, int rpV) -> int {   // destination of one column (hi tile byte offset): -1 none
            const int ty = code >> 13, ix = code & 0x1fff;
            if (ty == 0) return -1;
            if (ty == 1) return rpS + ((ix >> 2) << 7) + ((ix & 3) << 2);
            return A.oV + (ty - 2) * A.halfV + rpV + ((ix >> 2) << 7) + ((ix & 3) << 2);
        };
        int gofs[MAXB], dd[MAXB][4];
        bool act[MAXB], fast[MAXB];
#pragma unroll
        for (int i = 0; i < MAXB; ++i) {
            const int g = (warp >> 3) + NG * i;
            act[i] = false; fast[i] = false; gofs[i] = 0;
            dd[i][0] = dd[i][1] = dd[i][2] = dd[i][3] = -1;
            if (g < A.nblk) {
                const BlockE B = blks[g];
                const int ch = B.cb * 4 + cq;
                if (ch < B.nch) {
                    act[i] = true;
                    gofs[i] = 4 * ch;
                    const int kc = (int)((B.fast4 >> (8 * cq)) & 255u);
                    if (kc) {
                        fast[i] = true;
                        dd[i][0] = rowpartS + ((kc - 1) << 7);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) dd[i][j] = enc(ccode[B.cum + 4 * ch + j], rowpartS, rowpartV);
                    }
                }
            }
        }
        // narrow pieces: at most one task (32 four-byte pieces) per warp
        bool nact = false;
        int nrow = 0, ngofs = 0, ndd = -1;
        const NarrowE NW = A.nar[warp < A.nnar ? warp : 0];   // hoisted: no dynamic parameter indexing in the tile loop
        if (warp < A.nnar) {
            const NarrowE N = NW;
            const int p = N.t * 32 + lane;
            nrow = p / N.w;
            ngofs = p - nrow * N.w;
            nact = nrow < TM2;
            if (nact)
                ndd = enc(ccode[N.cum + ngofs], (((nrow >> 3) * KQ1) << 7) + ((nrow & 7) << 4),
                          (((nrow >> 3) * KQ2) << 7) + ((nrow & 7) << 4));
        }
        float4 R4[MAXB];
        float RN = 0.0f;
        int nidx[MAXB], nnidx = 0;
        auto load_idx = [&](int it) {   // index values of tile `it` of this CTA (identity rows: |rows| < 2^31, host-checked)
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * TM2;
            long long gr = row0 + wrow;
            if (gr > R - 1) gr = R - 1;
#pragma unroll
            for (int i = 0; i < MAXB; ++i) {
                nidx[i] = (int)gr;
                if (act[i]) {
                    const int32_t* ip = blks[(warp >> 3) + NG * i].idx;
                    if (ip) nidx[i] = ldgi_v(ip + gr);
                }
            }
            if (nact) {
                long long g2 = row0 + nrow;
                if (g2 > R - 1) g2 = R - 1;
                const int32_t* ip = NW.idx;
                nnidx = ip ? ldgi_v(ip + g2) : (int)g2;
            }
        };
        auto load_rows = [&]() {        // gather with the index values in nidx
#pragma unroll
            for (int i = 0; i < MAXB; ++i) {
                if (act[i]) {
                    const BlockE& B = blks[(warp >> 3) + NG * i];
                    R4[i] = ldg4_v(B.base + (long long)nidx[i] * B.ld + gofs[i]);
                }
            }
            if (nact) RN = ldg1_v(NW.base + (long long)nnidx * NW.ld + ngofs);
        };
        auto put = [&](unsigned char* aset, int d, float x) {
            if (d < 0) return;
            float hi, lo;
            split_tf32(x, hi, lo);
            *reinterpret_cast<float*>(aset + d) = hi;
            *reinterpret_cast<float*>(aset + d + A.HALF) = lo;
        };
        auto build = [&](int b) {
            unsigned char* aset = smraw + A.o_a + b * A.a_bytes;
#pragma unroll
            for (int i = 0; i < MAXB; ++i) {
                if (!act[i]) continue;
                const float4 v = R4[i];
                if (fast[i]) {
                    float4 h, l;
                    split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
                    *reinterpret_cast<float4*>(aset + dd[i][0]) = h;
                    *reinterpret_cast<float4*>(aset + dd[i][0] + A.HALF) = l;
                } else {
                    put(aset, dd[i][0], v.x);
                    put(aset, dd[i][1], v.y);
                    put(aset, dd[i][2], v.z);
                    put(aset, dd[i][3], v.w);
                }
            }
            if (nact) put(aset, ndd, RN);
        };
        // epilogue part 1: TMEM accumulators of tile `it` -> raw tile in shared memory (affine output layout; the l=0
        // norms are folded into the weights)
        const int e = warp & 3, jq = warp >> 2;
        float* otile = reinterpret_cast<float*>(smraw + A.o_out);
        float* ptile = reinterpret_cast<float*>(smraw + A.o_post);
        constexpr bool gate = GATE;
        const int dout = A.d_out, dpost = A.epi.d_post;
        const int drow = 16 * e + (lane & 15);
        const bool rowlane = lane < 16;
        float4 ypre = make_float4(0.f, 0.f, 0.f, 0.f), ypre2 = ypre;   // in2 of this thread's two epilogue rows
        const int fg = lane >> 2, fq = lane & 3;                        // 16x256b fragment: rows fg, fg + 8; columns 2 fq, 2 fq + 1
        int segpre = -1;
        int* sseg = reinterpret_cast<int*>(smraw + A.o_sseg);
        float4* hsm = reinterpret_cast<float4*>(smraw + A.o_hs);
        auto prefetch_y = [&](int it) {   // in2 row (and segment id) of this thread's epilogue row of tile `it`, consumed by drain(it)
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * TM2;
            long long gr = row0 + drow;
            if (gr > R - 1) gr = R - 1;
            {
                long long ga = row0 + 16 * e + fg, gb = ga + 8;
                if (ga > R - 1) ga = R - 1;
                if (gb > R - 1) gb = R - 1;
                ypre = ldg4_v(A.in2 + ga * 4);
                ypre2 = ldg4_v(A.in2 + gb * 4);
            }
            if (SEG && jq == 0) {
                if (rowlane) segpre = ldgi_v(A.seg_idx + gr);
                else if (e == 0 && lane == 16) segpre = row0 > 0 ? ldgi_v(A.seg_idx + row0 - 1) : -1;
                else if (e == 0 && lane == 17) segpre = row0 + TM2 < R ? ldgi_v(A.seg_idx + row0 + TM2) : -1;
            }
        };
        auto drain = [&](int it) {
            const int b = it & 1;
            // per row: Y0, c3 Y1x, c3 Y1y, c3 Y1z, c3 Y0
            const float ya[5] = {ypre.x, C3f * ypre.y, C3f * ypre.z, C3f * ypre.w, C3f * ypre.x};
            const float yb[5] = {ypre2.x, C3f * ypre2.y, C3f * ypre2.z, C3f * ypre2.w, C3f * ypre2.x};
            mbar_wait(BAR(2 + b), (it >> 1) & 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + (uint32_t)b * ACC + ((uint32_t)(32 * e) << 16);
            float* rowa = otile + (16 * e + fg) * A.dop;
            float* rowb = rowa + 8 * A.dop;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int cb = 8 * (jq + (W2 / 4) * h);   // 8-column blocks jq, jq + W2/4 (N1 <= 64: at most 8 blocks)
                if (cb < A.N1) {
                    float p[4], ux[4], uy[4], uz[4];
                    tc_ld_16x256(acc + cb, p);
                    tc_ld_16x256(acc + A.N1 + cb, ux);
                    tc_ld_16x256(acc + 2 * A.N1 + cb, uy);
                    tc_ld_16x256(acc + 3 * A.N1 + cb, uz);
                    tc_wait_ld();
                    if (cb < A.N2) {
                        const int m = cb + 2 * fq;              // z channels m, m + 1
                        float2 oa, ob;
                        oa.x = fmaf(ya[0], p[0], fmaf(ya[1], ux[0], fmaf(ya[2], uy[0], ya[3] * uz[0])));
                        oa.y = fmaf(ya[0], p[1], fmaf(ya[1], ux[1], fmaf(ya[2], uy[1], ya[3] * uz[1])));
                        ob.x = fmaf(yb[0], p[2], fmaf(yb[1], ux[2], fmaf(yb[2], uy[2], yb[3] * uz[2])));
                        ob.y = fmaf(yb[0], p[3], fmaf(yb[1], ux[3], fmaf(yb[2], uy[3], yb[3] * uz[3])));
                        if (m + 1 < A.mz) {
                            *reinterpret_cast<float2*>(rowa + A.oz0 + m) = oa;
                            *reinterpret_cast<float2*>(rowb + A.oz0 + m) = ob;
                        } else if (m < A.mz) {
                            rowa[A.oz0 + m] = oa.x;
                            rowb[A.oz0 + m] = ob.x;
                        }
                    } else {
                        const int m = cb - A.N2 + 2 * fq;       // v channels m, m + 1
                        const float* nv3 = norm + 3 * m;
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            if (m + u < A.mv) {
                                float* oa = rowa + A.ov0 + 3 * (m + u);
                                float* ob = rowb + A.ov0 + 3 * (m + u);
                                oa[0] = nv3[3 * u] * fmaf(ya[1], p[u], ya[4] * ux[u]);
                                oa[1] = nv3[3 * u + 1] * fmaf(ya[2], p[u], ya[4] * uy[u]);
                                oa[2] = nv3[3 * u + 2] * fmaf(ya[3], p[u], ya[4] * uz[u]);
                                ob[0] = nv3[3 * u] * fmaf(yb[1], p[2 + u], yb[4] * ux[2 + u]);
                                ob[1] = nv3[3 * u + 1] * fmaf(yb[2], p[2 + u], yb[4] * uy[2 + u]);
                                ob[2] = nv3[3 * u + 2] * fmaf(yb[3], p[2 + u], yb[4] * uz[2 + u]);
                            }
                        }
                    }
                }
            }
            if (SEG && jq == 0) {
                if (rowlane) sseg[1 + drow] = segpre;
                else if (e == 0 && lane == 16) sseg[0] = segpre;
                else if (e == 0 && lane == 17) sseg[65] = segpre;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(4 + b));
        };
        // epilogue part 2 (all worker threads): raw tile -> global (+residual); gate -> post rows / segment sum
        auto finish = [&](int it) {
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * TM2;
            const int nvalid = (int)min((long long)TM2, R - row0);
            const bool to_ptile = SEG && gate;
            if (A.out_raw) {
                float* dst = A.out_raw + row0 * dout;
                const float* res = A.resid ? A.resid + row0 * dout : nullptr;
                const int total = nvalid * dout;
                if (A.dop == dout) {   // the tile is the exact image of the global rows: linear 16-byte copy
                    const int n4 = total >> 2;
                    for (int t = tid; t < n4; t += WT) {
                        float4 v = reinterpret_cast<const float4*>(otile)[t];
                        if (res) {
                            const float4 q = __ldg(reinterpret_cast<const float4*>(res) + t);
                            v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
                        }
                        reinterpret_cast<float4*>(dst)[t] = v;
                    }
                    for (int t = (n4 << 2) + tid; t < total; t += WT) dst[t] = otile[t] + (res ? __ldg(res + t) : 0.0f);
                } else if ((dout & 1) == 0) {   // 8-byte pairs, rows at stride dop (even)
                    const int hw = dout >> 1, n2 = nvalid * hw;
                    for (int t = tid; t < n2; t += WT) {
                        const int r = (int)__umulhi((unsigned)t, A.mg_h), c2 = t - r * hw;
                        float2 v = *reinterpret_cast<const float2*>(otile + r * A.dop + 2 * c2);
                        if (res) {
                            const float2 q = __ldg(reinterpret_cast<const float2*>(res) + t);
                            v.x += q.x; v.y += q.y;
                        }
                        reinterpret_cast<float2*>(dst)[t] = v;
                    }
                } else {
                    for (int r = warp; r < nvalid; r += W2) {
                        const float* orow2 = otile + r * A.dop;
                        for (int c = lane; c < dout; c += 32) {
                            float v = orow2[c];
                            if (res) v += __ldg(res + r * dout + c);
                            dst[r * dout + c] = v;
                        }
                    }
                }
            }
            if (gate) {
                float* dstp = A.out_post ? A.out_post + row0 * dpost : nullptr;
                const int nsg = A.epi.ns_g;
                const int ns_tot = nvalid * nsg;
                for (int t = tid; t < ns_tot; t += WT) {          // swish on the scalars
                    const int r = (int)__umulhi((unsigned)t, A.mg_ns), c = t - r * nsg;
                    const float x = otile[r * A.dop + A.oz0 + c];
                    const float v = A.epi.cs * x * sigm(x);
                    if (dstp) dstp[r * dpost + c] = v;
                    if (to_ptile) ptile[r * A.dpp + c] = v;
                }
                const int nv_tot = nvalid * A.mv;
                for (int t = tid; t < nv_tot; t += WT) {          // sigmoid gate on the vectors, one channel per thread
                    const int r = (int)__umulhi((unsigned)t, A.mg_mv), v = t - r * A.mv;
                    const float* orow2 = otile + r * A.dop;
                    const float s = A.epi.cg * sigm(orow2[A.oz0 + nsg + v]);
                    const float* xv = orow2 + A.ov0 + 3 * v;
                    const float a0 = s * xv[0], a1 = s * xv[1], a2 = s * xv[2];
                    if (dstp) {
                        float* d = dstp + r * dpost + nsg + 3 * v;
                        d[0] = a0; d[1] = a1; d[2] = a2;
                    }
                    if (to_ptile) {
                        float* d = ptile + r * A.dpp + nsg + 3 * v;
                        d[0] = a0; d[1] = a1; d[2] = a2;
                    }
                }
            }
            if (SEG) {
                if (to_ptile) named_bar(1, WT);
                const float* stile = gate ? ptile : otile;
                const int sstr = gate ? A.dpp : A.dop, swid = gate ? dpost : dout;
                sorted_segment_sum_tile<WT>(stile, sstr, swid, nvalid, sseg, A.out_seg, swid, hsm, tid, 3);
            }
        };

        // ---- software pipeline: loads of tile it+1 are in flight while tile it is built and tile it-1 is finished
        if (nt > 0) {
            load_idx(0);
            load_rows();
            if (nt > 1) load_idx(1);
        }
        for (int it = 0; it < nt; ++it) {
            build(it & 1);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(it & 1));
            if (it + 1 < nt) {
                load_rows();
                if (it + 2 < nt) load_idx(it + 2);
            }
            if (it >= 1) {
                named_bar(2, WT);          // every worker is done reading the smem tile of tile it-2
                drain(it - 1);
                named_bar(1, WT);
                finish(it - 1);
            }
            prefetch_y(it);
        }
        if (nt > 0) {
            named_bar(2, WT);
            drain(nt - 1);
            named_bar(1, WT);
            finish(nt - 1);
        }
    }
    // ---------------- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == W2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

}  // namespace se3

using namespace se3;

// Launches the second-generation forward when the configuration is eligible (launched=false otherwise: the caller
// falls through to the generic fp32 kernel).
int se3_l1tp_tc2_try_forward(const int n[4], const int m[4], const int t_in[4], const int t_out[4], int ntab,
                             const int* h_tab, const int* d_tab, const se3_l1tp_fwd_args* a, const RowSrc& src,
                             const EpiL& epi, cudaStream_t st, bool* launched) {
    *launched = false;
    static int disabled = -1;
    if (disabled < 0) {
        const char* e = getenv("SE3_DISABLE_TC2");
        const char* e1 = getenv("SE3_DISABLE_TC");
        disabled = ((e && e[0] == '1') || (e1 && e1[0] == '1')) ? 1 : 0;
    }
    if (disabled) return SE3_OK;
    if (n[1] || n[2] || m[1] || m[2]) return SE3_OK;
    const int ns = n[0], nd = n[3], mz = m[0], mv = m[3];
    if (ns < 1 || nd < 1 || mz < 1 || mv < 1) return SE3_OK;
    if (a->rows >= (1ll << 31) - TM2) return SE3_OK;
    if (((uintptr_t)a->in2 & 15) != 0) return SE3_OK;
    const int dtot = src.cum[src.nseg];
    if (dtot > MAXCOL - 4) return SE3_OK;
    static thread_local Tc2Args A;   // large: keep off the stack; one per host thread (autograd runs backward on its own threads)
    memset(&A, 0, sizeof(A));
    A.ns = ns; A.nd = nd; A.mz = mz; A.mv = mv;
    A.N2 = (mz + 7) & ~7;
    A.N1 = A.N2 + ((mv + 7) & ~7);
    if (A.N1 > 64) return SE3_OK;   // 2 x 4 N1 TMEM columns
    // ---- column kinds of the concatenated row
    std::vector<int> kind(dtot, 0), chan(dtot, -1);   // 1 scalar, 2..4 vector x/y/z
    for (int k = 0; k < ns; ++k) { const int c = h_tab[t_in[0] + k]; kind[c] = 1; chan[c] = k; }
    for (int k = 0; k < nd; ++k) {
        const int c = h_tab[t_in[3] + k];
        for (int q = 0; q < 3; ++q) { kind[c + q] = 2 + q; chan[c + q] = k; }
    }
    // ---- scalar K slots: wide segments keep their 16-byte chunking, narrow segments fill the padding
    std::vector<int> slot(dtot, -1);
    std::vector<int> freeslots;
    int wide[SE3_MAX_SEG];
    int K1 = 0;
    for (int s = 0; s < src.nseg; ++s) {
        const int w = src.cum[s + 1] - src.cum[s];
        wide[s] = ((w & 3) == 0 && (src.ld[s] & 3) == 0 && ((uintptr_t)src.base[s] & 15) == 0) ? 1 : 0;
        if (!wide[s]) {
            if (w > 16) return SE3_OK;
            continue;
        }
        int first = -1, last = -1;
        for (int c = 0; c < w; ++c)
            if (kind[src.cum[s] + c] == 1) { if (first < 0) first = c; last = c; }
        if (first < 0) continue;
        const int lc0 = first & ~3, span = ((last + 4) & ~3) - lc0;
        for (int c = lc0; c < lc0 + span; ++c) {
            if (c < w && kind[src.cum[s] + c] == 1) slot[src.cum[s] + c] = K1 + c - lc0;
            else freeslots.push_back(K1 + c - lc0);
        }
        K1 += span;
    }
    size_t fp = 0;
    for (int s = 0; s < src.nseg; ++s) {
        if (wide[s]) continue;
        const int w = src.cum[s + 1] - src.cum[s];
        for (int c = 0; c < w; ++c)
            if (kind[src.cum[s] + c] == 1) slot[src.cum[s] + c] = fp < freeslots.size() ? freeslots[fp++] : K1++;
    }
    K1 = (K1 + 7) & ~7;
    const int K2 = (nd + 7) & ~7;
    if (K1 > MAXK1 || K1 < 8) return SE3_OK;
    A.K1 = K1; A.K2 = K2;
    for (int k = 0; k < MAXK1; ++k) A.sl2ch[k] = -1;
    for (int c = 0; c < dtot; ++c) {
        if (kind[c] == 1) { A.ccode[c] = (unsigned short)((1 << 13) | slot[c]); A.sl2ch[slot[c]] = (short)chan[c]; }
        else if (kind[c] >= 2) A.ccode[c] = (unsigned short)((kind[c] << 13) | chan[c]);
    }
    // ---- affine output layout (every irreps string of the form  c x0e + d x1o  in one block each)
    A.oz0 = h_tab[t_out[0]]; A.ov0 = h_tab[t_out[3]];
    for (int k = 0; k < mz; ++k) if (h_tab[t_out[0] + k] != A.oz0 + k) return SE3_OK;
    for (int k = 0; k < mv; ++k) if (h_tab[t_out[3] + k] != A.ov0 + 3 * k) return SE3_OK;
    if (epi.mode == SE3_EPI_GATE && (epi.ns_g < 1 || epi.ns_g + mv != mz)) return SE3_OK;
    // ---- wide blocks (warp w: row group w & 7, blocks (w >> 3) + NG i) and narrow tasks (one per warp at most)
    int nblk = 0, nnar = 0;
    for (int s = 0; s < src.nseg; ++s) {
        const int w = src.cum[s + 1] - src.cum[s];
        if (wide[s]) {
            const int nch = w >> 2, ncb = (nch + 3) >> 2;
            for (int cb = 0; cb < ncb; ++cb) {
                if (nblk >= NG * MAXB) return SE3_OK;
                BlockE& B = A.blk[nblk++];
                B.base = src.base[s]; B.idx = src.idx[s]; B.ld = src.ld[s]; B.cum = src.cum[s]; B.nch = nch; B.cb = cb;
                B.fast4 = 0; B.pad = 0;
                for (int q = 0; q < 4; ++q) {
                    const int ch = cb * 4 + q;
                    if (ch >= nch) continue;
                    const int c0 = src.cum[s] + 4 * ch;
                    const bool f = kind[c0] == 1 && kind[c0 + 1] == 1 && kind[c0 + 2] == 1 && kind[c0 + 3] == 1 &&
                                   (slot[c0] & 3) == 0 && slot[c0 + 1] == slot[c0] + 1 && slot[c0 + 2] == slot[c0] + 2 &&
                                   slot[c0 + 3] == slot[c0] + 3;
                    if (f) B.fast4 |= (unsigned)((slot[c0] >> 2) + 1) << (8 * q);
                }
            }
        } else {
            const int nt_ = (TM2 * w + 31) / 32;
            for (int t = 0; t < nt_; ++t) {
                if (nnar >= W2) return SE3_OK;
                NarrowE& N = A.nar[nnar++];
                N.base = src.base[s]; N.idx = src.idx[s]; N.ld = src.ld[s]; N.cum = src.cum[s]; N.w = w; N.t = t;
            }
        }
    }
    A.nblk = nblk; A.nnar = nnar;
    A.rows = a->rows; A.in2 = a->in2; A.wz = a->w[0]; A.wv = a->w[3]; A.nz = a->norm[0]; A.nv = a->norm[3];
    A.epi = epi; A.out_raw = a->out_raw; A.out_post = a->out_post; A.resid = a->resid; A.seg_idx = a->seg_idx;
    A.out_seg = a->out_seg;
    A.d_out = mz + 3 * mv;
    A.dop = ((A.d_out & 1) || (A.d_out & 3) == 2) ? A.d_out : A.d_out + 2;
    A.dpp = (epi.d_post & 3) == 0 ? tc_stage_stride(epi.d_post) : (epi.d_post | 1);
    if ((A.dop & 1) || (A.oz0 & 1)) return SE3_OK;   // the epilogue stores z outputs in 8-byte pairs
    auto magic = [](int d) { return d > 1 ? (unsigned)((0x100000000ull + (unsigned)d - 1) / (unsigned)d) : 0u; };
    if (epi.mode == SE3_EPI_GATE && (epi.ns_g < 2 || mv < 2)) return SE3_OK;   // magic division needs d >= 2
    A.mg_ns = magic(epi.ns_g); A.mg_mv = magic(mv); A.mg_h = magic(std::max(2, A.d_out >> 1));
    if ((A.d_out >> 1) < 2) return SE3_OK;
    auto al = [](int x, int q) { return (x + q - 1) / q * q; };
    A.halfS = TM2 * K1 * 4; A.halfV = TM2 * K2 * 4; A.oV = A.halfS;
    A.HALF = A.halfS + 3 * A.halfV;
    A.a_bytes = 2 * A.HALF;
    int o = 0;
    A.o_b1 = o; o += 2 * A.N1 * K1 * 4;
    A.o_b2 = o; o += 2 * A.N1 * K2 * 4;
    o = al(o, 1024);
    A.o_a = o; o += 2 * A.a_bytes;
    A.o_out = o; o += al(TM2 * A.dop * 4, 16);
    A.o_post = o; o += (epi.mode == SE3_EPI_GATE && a->seg_idx) ? al(TM2 * A.dpp * 4, 16) : 0;
    A.o_norm = o; o += al((3 * mv) * 4, 16);
    A.o_blk = o; o += al(NG * MAXB * (int)sizeof(BlockE), 16);
    A.o_ccode = o; o += al(MAXCOL * 2, 16);
    A.o_bar = o; o += 8 * 8 + 16;
    A.o_sseg = o; o += a->seg_idx ? 68 * 4 : 0;
    A.o_hs = o; o += a->seg_idx ? 8 * std::max(epi.d_post, A.d_out) * 16 : 0;
    int dev = 0, maxsm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (o > maxsm) return SE3_OK;
    const int smem = std::max(o, 120 * 1024);   // > half an SM: one CTA per SM owns all 512 TMEM columns
    static bool attr_set = false;
    if (!attr_set) {
        SE3_CUDA_TRY(cudaFuncSetAttribute(l1tp_tc2_fwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
        SE3_CUDA_TRY(cudaFuncSetAttribute(l1tp_tc2_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
        SE3_CUDA_TRY(cudaFuncSetAttribute(l1tp_tc2_fwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
        SE3_CUDA_TRY(cudaFuncSetAttribute(l1tp_tc2_fwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
        attr_set = true;
    }
    const long long ntiles = (a->rows + TM2 - 1) / TM2;
    const int grid = (int)std::min<long long>(ntiles, num_sms());
    const bool g = epi.mode == SE3_EPI_GATE;
    if (a->seg_idx) { if (g) l1tp_tc2_fwd_kernel<true, true><<<grid, T2_THREADS, smem, st>>>(A); else l1tp_tc2_fwd_kernel<true, false><<<grid, T2_THREADS, smem, st>>>(A); }
    else { if (g) l1tp_tc2_fwd_kernel<false, true><<<grid, T2_THREADS, smem, st>>>(A); else l1tp_tc2_fwd_kernel<false, false><<<grid, T2_THREADS, smem, st>>>(A); }
    SE3_LAUNCHED();
    g_tc_launches.fetch_add(1, std::memory_order_relaxed);
    *launched = true;
    return SE3_OK;
}
