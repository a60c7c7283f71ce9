// Fused forward of the SEGNN message layer on sm_100a (north-star kernel 4 + 5 in one launch):
//
//     gather node tables (x[dst], x[src] after the node-level weight contraction, csrc/msg_table.cu)
//       -> combine with SH(1), gate            = message 1  (never a GEMM operand in HBM)
//       -> weight contraction of message 2 on tcgen05 (3xTF32, fp32 accumulation in TMEM)
//       -> combine with SH(1), gate            = message 2
//       -> sorted-segment sum over the destination (no atomics inside a tile)  -> agg [Nn, ns + 3 nv]
//
// It replaces, per edge, two applications of L1TensorProduct.forward (L1TP:242-297) with the public-SEGNN swish gate in
// between and the scatter-add aggregation after them.  What still goes to HBM per edge is only what the backward reads:
// the two pre-activations and the gated message 1.
//
// Structure (the proven skeleton of l1tp_tc2.cu: one persistent CTA per SM, 16 homogeneous worker warps + 1 MMA warp,
// 64-row tiles, operand sets and TMEM accumulators double buffered, one mbarrier hand-off per tile), specialised at
// compile time for the hidden irreps NS x0e + NV x1o:
//   * build: lane = (row, unit of 4 table channels): 16-byte table loads (dst and src half), 4 FMAs per channel with
//     the SH, gate in registers, message-1 row written as tf32 hi/lo straight into the K-major UMMA operand tiles
//     S [64 x K1] and Vx, Vy, Vz [64 x K2] (the vector components are produced de-interleaved: no shuffle);
//   * MMA: P = S . B1, Uc = Vc . B2 with all norms and 1/sqrt(3) folded into B; the output COLUMNS are permuted so that
//     a gate scalar and the vector channel it gates are neighbours and land in the same thread of the 16x256b
//     accumulator fragment;
//   * drain: tcgen05.ld 16x256b -> 4 FMAs per channel with the SH -> gate in registers -> pre-activation tile and
//     message tile in shared memory;  finish: coalesced copy of the pre-activation, sorted-segment sum of the messages.
#include <algorithm>
#include <type_traits>

#include "tc_common.cuh"

namespace se3 {

static constexpr int FW = 16;                  // worker warps
static constexpr int F_THREADS = (FW + 1) * 32;
static constexpr int FWT = FW * 32;
static constexpr int FTM = 64;                 // rows per tile
static constexpr int NDMAX = 16;               // staged distinct destination rows per tile (more: read from global)

template <int NS, int NV>
struct FusedDims {
    static constexpr int MZ = NS + NV, CH = NS + 2 * NV, DPRE = NS + 4 * NV, D = NS + 3 * NV;
    static constexpr int HALF = 4 * CH, LDT = 2 * HALF;
    static constexpr int K1 = (NS + 7) & ~7, K2 = (NV + 7) & ~7, KQ1 = K1 / 4, KQ2 = K2 / 4;
    static constexpr int N = 64;                                   // accumulator columns (8 blocks of 8)
    static constexpr int FB = NS / 8, RS = NS % 8, PBF = NV / 4, RP = NV % 4;
    static constexpr int NBLK = FB + PBF + ((RS || RP) ? 1 : 0);   // used 8-column blocks
    static constexpr int SQ = NS / 4, RS4 = NS % 4, NP = NV / 2;   // build units per row: full scalar quads, partial, pairs
    static constexpr int NU = SQ + (RS4 ? 1 : 0) + NP;
    static constexpr int OSTR = DPRE;                              // pre-activation tile: exact image of the global rows
    static constexpr int PSTR = ((D + 3) & ~3) + 4;                // message tile stride (stride % 8 == 4)
    static constexpr int STGB = ((HALF * 4 - 112 + 127) / 128) * 128 + 112;  // bytes per staged row: >= HALF floats, 28 (mod 32) words
    static constexpr int halfS = FTM * K1 * 4, halfV = FTM * K2 * 4, HALFB = halfS + 3 * halfV, ABYTES = 2 * HALFB;
    static_assert(NS % 2 == 0 && NV % 2 == 0, "even channel counts (8-byte stores)");
    static_assert(D <= 64, "segment sum: one thread per column and run slot");
    static_assert(2 * RP + RS <= 8 && NBLK <= 8, "output columns do not fit 64 accumulator columns");
    static_assert(SQ <= 8 && NP + (RS4 ? 1 : 0) <= 8, "one round of scalar quads and one of pairs per warp pair");
    static_assert(STGB >= HALF * 4 && STGB % 16 == 0 && (HALF * 4) % 16 == 0, "staged rows are 16-byte multiples");
    static_assert((DPRE & 3) == 2 || (DPRE & 3) == 0, "pre-activation tile copy");
};

// output channel of accumulator column n: < MZ: 0e channel (scalar or gate), >= MZ: vector channel MZ + v, -1: unused
template <int NS, int NV>
__host__ __device__ constexpr int fused_colch(int n) {
    using F = FusedDims<NS, NV>;
    const int blk = n >> 3, j = n & 7;
    if (blk < F::FB) return 8 * blk + j;
    if (blk < F::FB + F::PBF) { const int v = 4 * (blk - F::FB) + (j >> 1); return (j & 1) ? F::MZ + v : NS + v; }
    if (blk == F::FB + F::PBF) {
        if (j < 2 * F::RP) { const int v = 4 * F::PBF + (j >> 1); return (j & 1) ? F::MZ + v : NS + v; }
        if (j < 2 * F::RP + F::RS) return 8 * F::FB + j - 2 * F::RP;
    }
    return -1;
}

// drain tasks: the used 8-column accumulator blocks spread over the four task slots jq (two blocks each at most) by
// longest-processing-time-first on their costs (tail block > gate/vector block > scalar block)
template <int NS, int NV>
struct DrainMap {
    int blk[4][2];
    constexpr DrainMap() : blk{{-1, -1}, {-1, -1}, {-1, -1}, {-1, -1}} {
        using F = FusedDims<NS, NV>;
        int load[4] = {0, 0, 0, 0}, cnt[4] = {0, 0, 0, 0};
        // blocks in decreasing cost: tail, pair blocks, scalar blocks
        int order[8] = {0, 0, 0, 0, 0, 0, 0, 0}, cost[8] = {0, 0, 0, 0, 0, 0, 0, 0}, n = 0;
        if (F::NBLK > F::FB + F::PBF) { order[n] = F::FB + F::PBF; cost[n] = 22; ++n; }
        for (int b = F::FB; b < F::FB + F::PBF; ++b) { order[n] = b; cost[n] = 16; ++n; }
        for (int b = 0; b < F::FB; ++b) { order[n] = b; cost[n] = 10; ++n; }
        for (int i = 0; i < n; ++i) {
            int best = -1;
            for (int q = 0; q < 4; ++q)
                if (cnt[q] < 2 && (best < 0 || load[q] < load[best])) best = q;
            blk[best][cnt[best]++] = order[i];
            load[best] += cost[i];
        }
    }
};

struct FusedFwdArgs {
    long long rows;            // edges
    const int* dst;            // [E] ascending
    const int* src;            // [E]
    const float* table;        // [n_all, LDT] node tables of message 1 (dst half | src half)
    const float* we;           // [2, CH] extras' weights of message 1
    const float* y;            // [E, 4]
    const float* extra;        // [E, 2]
    const float* wz2;          // message 2: weights_l0e [(NS + NV), MZ]
    const float* wv2;          // weights_l1o [(NS + NV), NV]
    const float* nz2;          // norm_l0e [MZ] or NULL
    const float* nv2;          // norm_l1o [3 NV] or NULL
    float* pre1;               // [E, DPRE]
    float* m1;                 // [E, D]
    float* pre2;               // [E, DPRE]
    float* agg;                // [n_dst, D], zero on entry
    float cs, cg;
    long long* dbg;            // NULL, or [grid][2][8] cycle counters of the phases (diagnostics: SE3_DBG_TIMING)
};

template <int NS, int NV>
struct FusedSmem {
    using F = FusedDims<NS, NV>;
    static constexpr int o_b1 = 0;
    static constexpr int o_b2 = o_b1 + 2 * F::N * F::K1 * 4;
    static constexpr int o_a = (o_b2 + 2 * F::N * F::K2 * 4 + 1023) & ~1023;
    static constexpr int o_out = o_a + F::ABYTES;            // ONE operand set: the MMAs of tile t are long complete when tile t+1 is built
    static constexpr int o_post = o_out + ((FTM * F::OSTR * 4 + 15) & ~15);
    static constexpr int o_we = o_post + FTM * F::PSTR * 4;
    static constexpr int o_bar = o_we + ((2 * F::CH * 4 + 15) & ~15);
    static constexpr int o_sseg = o_bar + 8 * 8 + 16;
    static constexpr int o_hs = o_sseg + 68 * 4;
    static constexpr int o_stg = (o_hs + 127) & ~127;   // src table rows of the next tile (cp.async.bulk)
    static constexpr int o_dstg = o_stg + FTM * F::STGB;                // its distinct dst table rows (first NDMAX)
    static constexpr int o_slot = o_dstg + NDMAX * F::STGB;             // row -> dst slot
    static constexpr int total = o_slot + FTM * 4 + 16;                 // + the overflow flag
};

__device__ __forceinline__ void fbulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fmbar_arrive_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(bar), "r"(bytes) : "memory");
}

template <int NS, int NV>
__global__ void __launch_bounds__(F_THREADS, 1) msg_fused_fwd_kernel(const __grid_constant__ FusedFwdArgs A) {
    using F = FusedDims<NS, NV>;
    using SM = FusedSmem<NS, NV>;
    constexpr int MZ = F::MZ, CH = F::CH, N = F::N, K1 = F::K1, K2 = F::K2, KQ1 = F::KQ1, KQ2 = F::KQ2;
    extern __shared__ __align__(1024) unsigned char smraw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + SM::o_bar);
    const uint32_t bar0 = smem_u32(bars);
    // barriers: 0,1 operand set full | 2,3 accumulator full | 4,5 accumulator empty | 6 staged src rows landed (cp.async
    //           of every worker thread) | 7 staged dst rows landed (TMA)
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    float* we_s = reinterpret_cast<float*>(smraw + SM::o_we);

    // ---------------- one-time setup
    for (int t = tid; t < 2 * CH; t += F_THREADS) we_s[t] = __ldg(A.we + t);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(BAR(i), FW);
            mbar_init(BAR(2 + i), 1);
            mbar_init(BAR(4 + i), FW);
        }
        mbar_init(BAR(6), FWT);          // one cp.async.mbarrier.arrive.noinc per worker thread
        mbar_init(BAR(7), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // zero both operand sets (K padding is never written again)
        float4* z = reinterpret_cast<float4*>(smraw + SM::o_a);
        for (int t = tid; t < F::ABYTES >> 4; t += F_THREADS) z[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    {   // message-2 weights -> canonical K-major B tiles (hi | lo), norms and c3 folded, columns permuted
        unsigned char* b1 = smraw + SM::o_b1;
        for (int t = tid; t < N * K1; t += F_THREADS) {
            const int n = t / K1, k = t - n * K1;
            const int ch = fused_colch<NS, NV>(n);
            float x = 0.0f;
            if (ch >= 0 && k < NS) {
                if (ch < MZ) x = __ldg(A.wz2 + k * MZ + ch) * (A.nz2 ? __ldg(A.nz2 + ch) : 1.0f);
                else x = __ldg(A.wv2 + k * NV + ch - MZ) * C3f * (A.nv2 ? __ldg(A.nv2 + 3 * (ch - MZ)) : 1.0f);
            }
            float hi, lo;
            split_tf32(x, hi, lo);
            const int o = canon_off(n, k, KQ1);
            *reinterpret_cast<float*>(b1 + o) = hi;
            *reinterpret_cast<float*>(b1 + N * K1 * 4 + o) = lo;
        }
        unsigned char* b2 = smraw + SM::o_b2;
        for (int t = tid; t < N * K2; t += F_THREADS) {
            const int n = t / K2, k = t - n * K2;
            const int ch = fused_colch<NS, NV>(n);
            float x = 0.0f;
            if (ch >= 0 && k < NV) {
                if (ch < MZ) x = __ldg(A.wz2 + (NS + k) * MZ + ch) * C3f * (A.nz2 ? __ldg(A.nz2 + ch) : 1.0f);
                else x = __ldg(A.wv2 + (NS + k) * NV + ch - MZ) * C3f * (A.nv2 ? __ldg(A.nv2 + 3 * (ch - MZ)) : 1.0f);
            }
            float hi, lo;
            split_tf32(x, hi, lo);
            const int o = canon_off(n, k, KQ2);
            *reinterpret_cast<float*>(b2 + o) = hi;
            *reinterpret_cast<float*>(b2 + N * K2 * 4 + o) = lo;
        }
    }
    fence_proxy_async();
    if (warp == FW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const long long R = A.rows;
    const long long ntiles = (R + FTM - 1) / FTM;
    const int nt = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    constexpr uint32_t ACC = 4u * N;

    if (warp == FW) {
        // ================= MMA issuer
        const uint32_t sb = smem_u32(smraw);
        const uint32_t idesc = make_idesc(N);
        constexpr uint32_t sboS = KQ1 * 128, sboV = KQ2 * 128;
        const uint64_t dB1h = make_desc(sb + SM::o_b1, sboS), dB1l = make_desc(sb + SM::o_b1 + N * K1 * 4, sboS);
        const uint64_t dB2h = make_desc(sb + SM::o_b2, sboV), dB2l = make_desc(sb + SM::o_b2 + N * K2 * 4, sboV);
        constexpr uint32_t vstep = ((uint32_t)F::halfV) >> 4;
        // dst halves of the table rows of the NEXT tile: one cp.async.bulk (TMA, no tensor map) per DISTINCT destination,
        // issued by this (otherwise idle) warp as soon as every worker has arrived from the build of the current tile, i.e.
        // has finished reading the staged rows.  (The 64 src rows are gathered by the workers with 16-byte cp.async: a
        // bulk request per 864-byte row was measured at ~100 cycles each, longer than the epilogue they should hide in.)
        const uint32_t stg_u32 = smem_u32(smraw + SM::o_stg);
        int pf_dst0 = 0, pf_dst1 = 0;
        int* sslot = reinterpret_cast<int*>(smraw + SM::o_slot);
        auto load_pf = [&](int it) {
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * FTM;
            long long g0 = row0 + lane, g1 = g0 + 32;
            if (g0 > R - 1) g0 = R - 1;
            if (g1 > R - 1) g1 = R - 1;
            pf_dst0 = ldgi_v(A.dst + g0);
            pf_dst1 = ldgi_v(A.dst + g1);
        };
        auto issue_pf = [&]() {
            // dst is ascending: slot of a row = number of changes of dst up to it; the first row of a slot copies the
            // dst half of that node's table row (once per node and tile instead of once per edge)
            const int up0 = __shfl_up_sync(0xffffffffu, pf_dst0, 1), last0 = __shfl_sync(0xffffffffu, pf_dst0, 31);
            const int up1 = __shfl_up_sync(0xffffffffu, pf_dst1, 1);
            const bool f0 = lane > 0 && pf_dst0 != up0;
            const bool f1 = pf_dst1 != (lane == 0 ? last0 : up1);
            const unsigned b0 = __ballot_sync(0xffffffffu, f0), b1 = __ballot_sync(0xffffffffu, f1);
            const unsigned le = 0xffffffffu >> (31 - lane);
            const int s0 = __popc(b0 & le), s1 = __popc(b0) + __popc(b1 & le);
            sslot[lane] = s0;
            sslot[lane + 32] = s1;
            const int ndist = __popc(b0) + __popc(b1) + 1;
            const int ncopy = min(ndist, NDMAX);
            if (lane == 0) sslot[FTM] = ndist > NDMAX ? 1 : 0;
            if (lane == 0) fmbar_arrive_tx(BAR(7), ncopy * F::HALF * 4);
            __syncwarp();
            const uint32_t dstg_u32 = stg_u32 + (SM::o_dstg - SM::o_stg);
            if ((lane == 0 || f0) && s0 < NDMAX) fbulk_g2s(dstg_u32 + s0 * F::STGB, A.table + (long long)pf_dst0 * F::LDT, F::HALF * 4, BAR(7));
            if (f1 && s1 < NDMAX) fbulk_g2s(dstg_u32 + s1 * F::STGB, A.table + (long long)pf_dst1 * F::LDT, F::HALF * 4, BAR(7));
        };
        if (nt > 0) {
            load_pf(0);
            issue_pf();
            if (nt > 1) load_pf(1);
        }
        for (int it = 0; it < nt; ++it) {
            const int b = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            mbar_wait(BAR(b), ph);
            if (it + 1 < nt) {             // staging is free: every worker arrived from build(it); copies first, MMAs after
                issue_pf();
                if (it + 2 < nt) load_pf(it + 2);
            }
            mbar_wait(BAR(4 + b), ph ^ 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t acc = tmem_base + (uint32_t)b * ACC;
                const uint32_t aS = sb + SM::o_a;
                const uint64_t dSh = make_desc(aS, sboS), dSl = make_desc(aS + F::HALFB, sboS);
#pragma unroll
                for (int j = 0; j < K1 / 8; ++j) {
                    const uint64_t o = (uint64_t)(j * 16);
                    tc_mma_tf32(acc, dSh + o, dB1h + o, idesc, j ? 1u : 0u);
                    tc_mma_tf32(acc, dSh + o, dB1l + o, idesc, 1u);
                    tc_mma_tf32(acc, dSl + o, dB1h + o, idesc, 1u);
                }
                const uint64_t dVh = make_desc(aS + F::halfS, sboV), dVl = make_desc(aS + F::halfS + F::HALFB, sboV);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const uint32_t accc = acc + (uint32_t)(c + 1) * N;
                    const uint64_t co = (uint64_t)c * vstep;
#pragma unroll
                    for (int j = 0; j < K2 / 8; ++j) {
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
        const int* sslot = reinterpret_cast<const int*>(smraw + SM::o_slot);
        // build mapping: warp w owns row group rb = w & 7 (rows 8 rb + (lane & 7)); its lanes' units are
        // u = 8 (w >> 3) + 4 round + (lane >> 3), round = 0, 1
        const int r8 = lane & 7, cq = lane >> 3;
        const int wrow = (warp & 7) * 8 + r8;
        const int sub = warp >> 3;
        const int rowoffS = (((wrow >> 3) * KQ1) << 7) + ((wrow & 7) << 4);
        const int rowoffV = (((wrow >> 3) * KQ2) << 7) + ((wrow & 7) << 4);
        // row data of the tile being built, fetched one tile ahead
        int n_dst = 0;
        float4 n_y = make_float4(0.f, 0.f, 0.f, 0.f);
        float2 n_ex = make_float2(0.f, 0.f);
        auto load_row = [&](int it) {
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * FTM;
            long long gr = row0 + wrow;
            if (gr > R - 1) gr = R - 1;
            n_dst = ldgi_v(A.dst + gr);
            n_y = ldg4_v(A.y + 4 * gr);
            n_ex = ldg2_v(A.extra + 2 * gr);
        };
        // src halves of the table rows of tile `it` -> staging, 16 bytes per cp.async: eight threads per row (its index
        // is fetched a whole build ahead into ONE register), each thread copies every eighth piece -> 128 contiguous
        // bytes per row and step; completion on BAR(6) (every worker thread arrives once per tile)
        constexpr int PCS = F::HALF / 4;                         // 16-byte pieces per row
        static_assert(FWT == 8 * FTM, "eight gather threads per row");
        int g_src = 0;
        auto load_gsrc = [&](int it) {
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * FTM;
            long long gr = row0 + (tid >> 3);
            if (gr > R - 1) gr = R - 1;
            g_src = ldgi_v(A.src + gr);
        };
        auto gather_src = [&]() {
            const uint32_t dstp = smem_u32(smraw + SM::o_stg) + (tid >> 3) * F::STGB + 16 * (tid & 7);
            const float* srcp = A.table + (long long)g_src * F::LDT + F::HALF + 4 * (tid & 7);
#pragma unroll
            for (int j = 0; j < (PCS + 7) / 8; ++j)
                if (8 * j + (tid & 7) < PCS) cp_async16(dstp + 128 * j, srcp + 32 * j, true);
            cp_async_mbar_arrive_noinc(BAR(6));
        };
        auto st_hl4 = [&](unsigned char* p, float a, float b, float c, float d) {   // 16-byte operand piece, hi and lo
            float4 h, l;
            split_tf32(a, h.x, l.x); split_tf32(b, h.y, l.y); split_tf32(c, h.z, l.z); split_tf32(d, h.w, l.w);
            *reinterpret_cast<float4*>(p) = h;
            *reinterpret_cast<float4*>(p + F::HALFB) = l;
        };
        auto st_hl2 = [&](unsigned char* p, float a, float b) {
            float2 h, l;
            split_tf32(a, h.x, l.x); split_tf32(b, h.y, l.y);
            *reinterpret_cast<float2*>(p) = h;
            *reinterpret_cast<float2*>(p + F::HALFB) = l;
        };
        // combination of one table channel with the SH: P += extras, x = Y0 P + Y1 . U
        auto zval = [&](const float4& a, const float4& b, int ch, const float4& y, const float2& ex) {
            const float P = a.x + b.x + fmaf(ex.x, we_s[ch], ex.y * we_s[CH + ch]);
            return fmaf(y.x, P, fmaf(y.y, a.y + b.y, fmaf(y.z, a.z + b.z, y.w * (a.w + b.w))));
        };
        // OVF: the tile has more than NDMAX distinct destinations (rare): the dst halves of the rows past the staged slots
        // come straight from global memory through a generic pointer; in the common case every table load is an LDS
        auto build_body = [&](int it, auto OVF) {
            unsigned char* aset = smraw + SM::o_a;
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * FTM;
            const long long gr = row0 + wrow;
            const bool valid = gr < R;
            const float4 y = n_y;
            const float2 ex = n_ex;
            const float* ts = reinterpret_cast<const float*>(smraw + SM::o_stg + wrow * F::STGB);   // staged src half
            auto lds4 = [&](const float* q) { return *reinterpret_cast<const float4*>(q); };
            const int slot = sslot[wrow];
            const float* tds = reinterpret_cast<const float*>(smraw + SM::o_dstg + (slot < NDMAX ? slot : 0) * F::STGB);
            const float* td = (decltype(OVF)::value && slot >= NDMAX) ? A.table + (long long)n_dst * F::LDT : tds;
            float* pre = A.pre1 + gr * F::DPRE;
            float* m1 = A.m1 + gr * F::D;
#pragma unroll
            for (int round = 0; round < 2; ++round) {
                // round 0: scalar quads 4 sub + cq; round 1: pairs 4 sub + cq, then the partial quad (the heavy pair
                // units are spread over all warps: a warp's round costs the sum of the unit kinds its lanes hold)
                const int t = 4 * sub + cq;
                int u;
                if (round == 0) u = t < F::SQ ? t : F::NU;
                else u = t < F::NP ? F::SQ + (F::RS4 ? 1 : 0) + t : ((F::RS4 && t == F::NP) ? F::SQ : F::NU);
                if (u >= F::NU) continue;
                if (u < F::SQ) {
                    // four scalar channels 4u .. 4u+3
                    float4 a[4], b[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) { a[j] = lds4(td + 16 * u + 4 * j); b[j] = lds4(ts + 16 * u + 4 * j); }
                    float x[4], m[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) { x[j] = zval(a[j], b[j], 4 * u + j, y, ex); m[j] = A.cs * x[j] * sigm(x[j]); }
                    if (valid) {
                        *reinterpret_cast<float2*>(pre + 4 * u) = make_float2(x[0], x[1]);
                        *reinterpret_cast<float2*>(pre + 4 * u + 2) = make_float2(x[2], x[3]);
                        *reinterpret_cast<float4*>(m1 + 4 * u) = make_float4(m[0], m[1], m[2], m[3]);
                    }
                    st_hl4(aset + rowoffS + (u << 7), m[0], m[1], m[2], m[3]);
                } else if (F::RS4 && u == F::SQ) {
                    // the last RS4 (= 2) scalar channels
                    const float4 a0 = lds4(td + 16 * u), b0 = lds4(ts + 16 * u);
                    const float4 a1 = lds4(td + 16 * u + 4), b1 = lds4(ts + 16 * u + 4);
                    const float x0 = zval(a0, b0, 4 * u, y, ex), x1 = zval(a1, b1, 4 * u + 1, y, ex);
                    const float m0 = A.cs * x0 * sigm(x0), mm1 = A.cs * x1 * sigm(x1);
                    if (valid) {
                        *reinterpret_cast<float2*>(pre + 4 * u) = make_float2(x0, x1);
                        *reinterpret_cast<float2*>(m1 + 4 * u) = make_float2(m0, mm1);
                    }
                    st_hl4(aset + rowoffS + (u << 7), m0, mm1, 0.0f, 0.0f);
                } else {
                    // pair unit i: gates NS + 2i, NS + 2i + 1 and the vector channels they gate
                    const int i = u - F::SQ - (F::RS4 ? 1 : 0);
                    const int cg0 = NS + 2 * i, cv0 = MZ + 2 * i;
                    const float4 ga0 = lds4(td + 4 * cg0), gb0 = lds4(ts + 4 * cg0);
                    const float4 ga1 = lds4(td + 4 * cg0 + 4), gb1 = lds4(ts + 4 * cg0 + 4);
                    const float4 va0 = lds4(td + 4 * cv0), vb0 = lds4(ts + 4 * cv0);
                    const float4 va1 = lds4(td + 4 * cv0 + 4), vb1 = lds4(ts + 4 * cv0 + 4);
                    const float xg0 = zval(ga0, gb0, cg0, y, ex), xg1 = zval(ga1, gb1, cg0 + 1, y, ex);
                    const float P0 = va0.x + vb0.x + fmaf(ex.x, we_s[cv0], ex.y * we_s[CH + cv0]);
                    const float P1 = va1.x + vb1.x + fmaf(ex.x, we_s[cv0 + 1], ex.y * we_s[CH + cv0 + 1]);
                    const float p00 = fmaf(y.y, P0, y.x * (va0.y + vb0.y)), p01 = fmaf(y.z, P0, y.x * (va0.z + vb0.z)),
                                p02 = fmaf(y.w, P0, y.x * (va0.w + vb0.w));
                    const float p10 = fmaf(y.y, P1, y.x * (va1.y + vb1.y)), p11 = fmaf(y.z, P1, y.x * (va1.z + vb1.z)),
                                p12 = fmaf(y.w, P1, y.x * (va1.w + vb1.w));
                    const float s0 = A.cg * sigm(xg0), s1 = A.cg * sigm(xg1);
                    const float q00 = s0 * p00, q01 = s0 * p01, q02 = s0 * p02, q10 = s1 * p10, q11 = s1 * p11, q12 = s1 * p12;
                    if (valid) {
                        *reinterpret_cast<float2*>(pre + cg0) = make_float2(xg0, xg1);
                        float* pv = pre + MZ + 6 * i;
                        *reinterpret_cast<float2*>(pv) = make_float2(p00, p01);
                        *reinterpret_cast<float2*>(pv + 2) = make_float2(p02, p10);
                        *reinterpret_cast<float2*>(pv + 4) = make_float2(p11, p12);
                        float* qv = m1 + NS + 6 * i;
                        *reinterpret_cast<float2*>(qv) = make_float2(q00, q01);
                        *reinterpret_cast<float2*>(qv + 2) = make_float2(q02, q10);
                        *reinterpret_cast<float2*>(qv + 4) = make_float2(q11, q12);
                    }
                    unsigned char* v0 = aset + F::halfS + rowoffV + ((i >> 1) << 7) + ((i & 1) << 3);
                    st_hl2(v0, q00, q10);
                    st_hl2(v0 + F::halfV, q01, q11);
                    st_hl2(v0 + 2 * F::halfV, q02, q12);
                }
            }
        };
        auto build = [&](int it) {
            mbar_wait(BAR(6), (uint32_t)(it & 1));
            mbar_wait(BAR(7), (uint32_t)(it & 1));
            if (it >= 1) mbar_wait(BAR(2 + ((it - 1) & 1)), (uint32_t)(((it - 1) >> 1) & 1));   // MMAs of tile it-1 have read the set
            if (sslot[FTM] == 0) build_body(it, std::false_type{});
            else build_body(it, std::true_type{});
        };
        // ---- epilogue
        const int e = warp & 3, jq = warp >> 2;
        float* otile = reinterpret_cast<float*>(smraw + SM::o_out);
        float* ptile = reinterpret_cast<float*>(smraw + SM::o_post);
        const int fg = lane >> 2, fq = lane & 3;
        const int drow = 16 * e + (lane & 15);
        const bool rowlane = lane < 16;
        float4 ypre = make_float4(0.f, 0.f, 0.f, 0.f), ypre2 = ypre;
        int segpre = -1;
        int* sseg = reinterpret_cast<int*>(smraw + SM::o_sseg);
        auto prefetch_y = [&](int it) {   // SH rows (and segment ids) of this thread's epilogue rows of tile `it`
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * FTM;
            long long ga = row0 + 16 * e + fg, gb = ga + 8;
            if (ga > R - 1) ga = R - 1;
            if (gb > R - 1) gb = R - 1;
            ypre = ldg4_v(A.y + ga * 4);
            ypre2 = ldg4_v(A.y + gb * 4);
            if (jq == 0) {
                long long gr = row0 + drow;
                if (gr > R - 1) gr = R - 1;
                if (rowlane) segpre = ldgi_v(A.dst + gr);
                else if (e == 0 && lane == 16) segpre = row0 > 0 ? ldgi_v(A.dst + row0 - 1) : -1;
                else if (e == 0 && lane == 17) segpre = row0 + FTM < R ? ldgi_v(A.dst + row0 + FTM) : -1;
            }
        };
        auto zacc = [&](const float4& y, float p, float ux, float uy, float uz) {
            return fmaf(y.x, p, fmaf(y.y, ux, fmaf(y.z, uy, y.w * uz)));
        };
        // one (gate, vector) column pair of one row: pre-activation + gated message into the tiles
        auto pair_out = [&](float* orow, float* prow, int v, const float4& y, float pg, float uxg, float uyg, float uzg,
                            float pv, float uxv, float uyv, float uzv) {
            const float xg = zacc(y, pg, uxg, uyg, uzg);
            const float c0 = fmaf(y.y, pv, y.x * uxv), c1 = fmaf(y.z, pv, y.x * uyv), c2 = fmaf(y.w, pv, y.x * uzv);
            orow[NS + v] = xg;
            float* ov = orow + MZ + 3 * v;
            ov[0] = c0; ov[1] = c1; ov[2] = c2;
            const float s = A.cg * sigm(xg);
            float* pvv = prow + NS + 3 * v;
            pvv[0] = s * c0; pvv[1] = s * c1; pvv[2] = s * c2;
        };
        auto scal_out = [&](float* orow, float* prow, int s0, const float4& y, float p0, float ux0, float uy0, float uz0,
                            float p1, float ux1, float uy1, float uz1) {
            const float x0 = zacc(y, p0, ux0, uy0, uz0), x1 = zacc(y, p1, ux1, uy1, uz1);
            *reinterpret_cast<float2*>(orow + s0) = make_float2(x0, x1);
            *reinterpret_cast<float2*>(prow + s0) = make_float2(A.cs * x0 * sigm(x0), A.cs * x1 * sigm(x1));
        };
        auto drain = [&](int it) {
            const int b = it & 1;
            const float4 ya = ypre, yb = ypre2;
            mbar_wait(BAR(2 + b), (it >> 1) & 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + (uint32_t)b * ACC + ((uint32_t)(32 * e) << 16);
            float* oa = otile + (16 * e + fg) * F::OSTR;
            float* ob = oa + 8 * F::OSTR;
            float* pa = ptile + (16 * e + fg) * F::PSTR;
            float* pb = pa + 8 * F::PSTR;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                constexpr DrainMap<NS, NV> DM{};
                const int blk = h == 0 ? (jq == 0 ? DM.blk[0][0] : (jq == 1 ? DM.blk[1][0] : (jq == 2 ? DM.blk[2][0] : DM.blk[3][0])))
                                       : (jq == 0 ? DM.blk[0][1] : (jq == 1 ? DM.blk[1][1] : (jq == 2 ? DM.blk[2][1] : DM.blk[3][1])));
                if (blk >= 0) {
                    float p[4], ux[4], uy[4], uz[4];
                    const int cb = 8 * blk;
                    tc_ld_16x256(acc + cb, p);
                    tc_ld_16x256(acc + N + cb, ux);
                    tc_ld_16x256(acc + 2 * N + cb, uy);
                    tc_ld_16x256(acc + 3 * N + cb, uz);
                    tc_wait_ld();
                    if (blk < F::FB) {
                        const int s0 = cb + 2 * fq;
                        scal_out(oa, pa, s0, ya, p[0], ux[0], uy[0], uz[0], p[1], ux[1], uy[1], uz[1]);
                        scal_out(ob, pb, s0, yb, p[2], ux[2], uy[2], uz[2], p[3], ux[3], uy[3], uz[3]);
                    } else if (blk < F::FB + F::PBF) {
                        const int v = 4 * (blk - F::FB) + fq;
                        pair_out(oa, pa, v, ya, p[0], ux[0], uy[0], uz[0], p[1], ux[1], uy[1], uz[1]);
                        pair_out(ob, pb, v, yb, p[2], ux[2], uy[2], uz[2], p[3], ux[3], uy[3], uz[3]);
                    } else {
                        if (fq < F::RP) {
                            const int v = 4 * F::PBF + fq;
                            pair_out(oa, pa, v, ya, p[0], ux[0], uy[0], uz[0], p[1], ux[1], uy[1], uz[1]);
                            pair_out(ob, pb, v, yb, p[2], ux[2], uy[2], uz[2], p[3], ux[3], uy[3], uz[3]);
                        } else if (2 * (fq - F::RP) < F::RS) {
                            const int s0 = 8 * F::FB + 2 * (fq - F::RP);
                            scal_out(oa, pa, s0, ya, p[0], ux[0], uy[0], uz[0], p[1], ux[1], uy[1], uz[1]);
                            scal_out(ob, pb, s0, yb, p[2], ux[2], uy[2], uz[2], p[3], ux[3], uy[3], uz[3]);
                        }
                    }
                }
            }
            if (jq == 0) {
                if (rowlane) sseg[1 + drow] = segpre;
                else if (e == 0 && lane == 16) sseg[0] = segpre;
                else if (e == 0 && lane == 17) sseg[65] = segpre;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(4 + b));
        };
        auto finish = [&](int it) {
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * FTM;
            const int nvalid = (int)min((long long)FTM, R - row0);
            if (warp >= FW / 2) {
                // upper half of the workers: pre-activation of message 2 (the tile is the exact image of the global rows)
                float* dstp = A.pre2 + row0 * F::DPRE;
                const int total = nvalid * F::DPRE, n4 = total >> 2;
                for (int t = tid - FWT / 2; t < n4; t += FWT / 2) reinterpret_cast<float4*>(dstp)[t] = reinterpret_cast<const float4*>(otile)[t];
                for (int t = (n4 << 2) + tid - FWT / 2; t < total; t += FWT / 2) dstp[t] = otile[t];
            } else {
                // lower half: sorted-segment sum (north-star kernel 5), run based: the rows are sorted by destination, so a
                // tile is a handful of runs; every warp derives the run starts with two ballots, thread = (column, run
                // slot) sums whole runs, one writer per (node, column): plain store, except for the (at most two) runs
                // that continue in the neighbouring tiles, which use red.add.  No barrier, no merge pass.
                const int c = tid & 63, slot = tid >> 6;
                const bool s0 = lane < nvalid && (lane == 0 || sseg[1 + lane] != sseg[lane]);
                const bool s1 = 32 + lane < nvalid && sseg[33 + lane] != sseg[32 + lane];
                unsigned long long mask = (unsigned long long)__ballot_sync(0xffffffffu, s0) |
                                          ((unsigned long long)__ballot_sync(0xffffffffu, s1) << 32);
                const int prevseg = sseg[0], nextseg = sseg[65];
                int j = 0;
                while (mask) {
                    const int start = __ffsll((long long)mask) - 1;
                    mask &= mask - 1;
                    const int end = mask ? __ffsll((long long)mask) - 1 : nvalid;
                    if ((j & (FWT / 128 - 1)) == slot && c < F::D) {
                        const float* col = ptile + c;
                        float acc = 0.0f;
#pragma unroll 4
                        for (int r = start; r < end; ++r) acc += col[r * F::PSTR];
                        const int seg = sseg[1 + start];
                        float* o = A.agg + (long long)seg * F::D + c;
                        if ((start == 0 && seg == prevseg) || (end == nvalid && seg == nextseg)) atomicAdd(o, acc);
                        else *o = acc;
                    }
                    ++j;
                }
            }
        };

        if (nt > 0) {
            load_row(0);
            load_gsrc(0);
            gather_src();
        }
        long long tacc[6] = {0, 0, 0, 0, 0, 0};
        const bool timing = A.dbg != nullptr && lane == 0 && (warp == 0 || warp == 9);
        for (int it = 0; it < nt; ++it) {
            long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0;
            if (timing) t0 = clock64();
            if (it + 1 < nt) load_gsrc(it + 1);     // consumed after the build, by gather_src
            build(it);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(it & 1));
            if (it + 1 < nt) load_row(it + 1);
            if (timing) t1 = clock64();
            named_bar(2, FWT);             // every worker is done with the staged rows of tile it and the tiles of tile it-2
            if (it + 1 < nt) gather_src();
            if (timing) t2 = clock64();
            if (it >= 1) {
                drain(it - 1);
                if (timing) t3 = clock64();
                named_bar(1, FWT);
                if (timing) t4 = clock64();
                finish(it - 1);
            }
            prefetch_y(it);
            if (timing && it >= 1) {
                t5 = clock64();
                tacc[0] += t1 - t0; tacc[1] += t2 - t1; tacc[2] += t3 - t2; tacc[3] += t4 - t3; tacc[4] += t5 - t4; tacc[5] += 1;
            }
        }
        if (timing) {
            long long* o = A.dbg + ((long long)blockIdx.x * 2 + (warp == 0 ? 0 : 1)) * 8;
            for (int i = 0; i < 6; ++i) o[i] = tacc[i];
        }
        if (nt > 0) {
            named_bar(2, FWT);
            drain(nt - 1);
            named_bar(1, FWT);
            finish(nt - 1);
        }
    }
    // ---------------- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == FW) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

template <int NS, int NV>
static int msg_fused_fwd_launch(const FusedFwdArgs& A, cudaStream_t st) {
    using SM = FusedSmem<NS, NV>;
    int dev = 0, maxsm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (SM::total > maxsm) { set_error("msg_fused_forward: %d bytes of shared memory needed, %d available", SM::total, maxsm); return SE3_ERR_TOO_LARGE; }
    const int smem = std::max(SM::total, 120 * 1024);   // > half an SM: one CTA per SM owns all 512 TMEM columns
    static bool attr_set = false;
    if (!attr_set) {
        SE3_CUDA_TRY(cudaFuncSetAttribute(msg_fused_fwd_kernel<NS, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
        attr_set = true;
    }
    const long long ntiles = (A.rows + FTM - 1) / FTM;
    const int grid = (int)std::min<long long>(ntiles, num_sms());
    msg_fused_fwd_kernel<NS, NV><<<grid, F_THREADS, smem, st>>>(A);
    SE3_LAUNCHED();
    g_tc_launches.fetch_add(1, std::memory_order_relaxed);
    return SE3_OK;
}

}  // namespace se3

using namespace se3;

extern "C" int se3_msg_fused_forward_dbg(int32_t ns, int32_t nv, int64_t rows, const int32_t* dst, const int32_t* src,
                                         const float* table, const float* we, const float* y, const float* extra,
                                         const float* wz2, const float* wv2, const float* nz2, const float* nv2,
                                         float gate_cs, float gate_cg, float* pre1, float* m1, float* pre2, float* agg,
                                         int64_t* dbg, void* stream);

extern "C" int se3_msg_fused_supported(int32_t ns, int32_t nv, int32_t ne) {
    return ne == 2 && ((ns == 34 && nv == 10) || (ns == 16 && nv == 8)) ? 1 : 0;
}

extern "C" int se3_msg_fused_forward(int32_t ns, int32_t nv, int64_t rows, const int32_t* dst, const int32_t* src,
                                     const float* table, const float* we, const float* y, const float* extra,
                                     const float* wz2, const float* wv2, const float* nz2, const float* nv2, float gate_cs,
                                     float gate_cg, float* pre1, float* m1, float* pre2, float* agg, void* stream) {
    return se3_msg_fused_forward_dbg(ns, nv, rows, dst, src, table, we, y, extra, wz2, wv2, nz2, nv2, gate_cs, gate_cg, pre1, m1,
                                     pre2, agg, nullptr, stream);
}

// diagnostics: dbg = NULL or [148][2][8] int64 receiving the accumulated clock64() cycles of the phases of two warps
extern "C" int se3_msg_fused_forward_dbg(int32_t ns, int32_t nv, int64_t rows, const int32_t* dst, const int32_t* src,
                                         const float* table, const float* we, const float* y, const float* extra,
                                         const float* wz2, const float* wv2, const float* nz2, const float* nv2,
                                         float gate_cs, float gate_cg, float* pre1, float* m1, float* pre2, float* agg,
                                         int64_t* dbg, void* stream) {
    if (rows < 0 || rows >= (1ll << 31) - FTM) { set_error("msg_fused_forward: bad row count"); return SE3_ERR_INVALID; }
    if (rows == 0) return SE3_OK;
    if (!dst || !src || !table || !we || !y || !extra || !wz2 || !wv2 || !pre1 || !m1 || !pre2 || !agg) {
        set_error("msg_fused_forward: null argument");
        return SE3_ERR_INVALID;
    }
    if (((uintptr_t)table | (uintptr_t)y | (uintptr_t)m1 | (uintptr_t)pre2) & 15) { set_error("msg_fused_forward: 16-byte alignment"); return SE3_ERR_INVALID; }
    FusedFwdArgs A;
    A.rows = rows; A.dst = dst; A.src = src; A.table = table; A.we = we; A.y = y; A.extra = extra; A.wz2 = wz2; A.wv2 = wv2;
    A.nz2 = nz2; A.nv2 = nv2; A.pre1 = pre1; A.m1 = m1; A.pre2 = pre2; A.agg = agg; A.cs = gate_cs; A.cg = gate_cg;
    A.dbg = (long long*)dbg;
    if (ns == 34 && nv == 10) return msg_fused_fwd_launch<34, 10>(A, (cudaStream_t)stream);
    if (ns == 16 && nv == 8) return msg_fused_fwd_launch<16, 8>(A, (cudaStream_t)stream);
    set_error("msg_fused_forward: hidden irreps %dx0e+%dx1o are not instantiated", (int)ns, (int)nv);
    return SE3_ERR_INVALID;
}
