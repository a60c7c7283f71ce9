// Fused input-gradient kernel of the SEGNN message layer on sm_100a: everything between the cotangent of the
// aggregate and the cotangent of message 1's PRE-activation, per edge, in one launch:
//
//     g_agg[dst]  -> gate VJP of message 2 (saved pre-activation)          = g_pre2   (optionally written: the weight-
//                                                                                       gradient kernel reads it)
//                 -> weight contraction with W2^T on tcgen05 (3xTF32)      = cotangent of the gated message 1
//                 -> gate VJP of message 1 (saved pre-activation)          = g_pre1   [E, ns + 4 nv]
//
// g_pre1 then goes through the two segment-sum passes of csrc/msg_table.cu (MODE 2 / 1) and the node-level kernels of
// csrc/msg_node.cu.  It replaces the autograd backward of L1TensorProduct.forward (L1TP:242-297) for message 2's input
// and both gate backward passes (pure autograd in the reference).
//
// Math (norms and 1/sqrt(3) folded into the B tiles; g0 / g1 = scalar / vector part of g_pre2, S / V = message 1):
//     HZ = g0 [mz],  HG[m] = sum_c Y1c g1[m][c],  HVc[m] = g1[m][c]
//     GS1 = HZ . BZs,  GS2 = HG . BVs,  GD = HZ . BZd,  GTc = HVc . BVv          (42 MMAs M=64 per tile: [GS1 | GD] is
//     one product, the two B tiles share the operand HZ)
//     gS_k = Y0 GS1[k] + GS2[k]          gV_kc = Y1c GD[k] + Y0 GTc[k]
// Same skeleton as msg_fused_fwd.cu: persistent CTA per SM, 16 worker warps + 1 MMA warp, 64-row tiles, operand sets
// and accumulators double buffered.  Build = streaming loads (the cotangent rows are gathered through the SORTED dst
// index: L1 hits) + gate VJP in registers + tf32 hi/lo K-major operand stores; drain = tcgen05.ld 16x256b, SH combine,
// gate VJP with the prefetched pre-activation of message 1, result tile in shared memory; finish = coalesced copy-out.
#include <algorithm>
#include <type_traits>

#include "tc_common.cuh"

namespace se3 {

static constexpr int BWK = 16;                 // worker warps
static constexpr int B_THREADS = (BWK + 1) * 32;
static constexpr int BWT = BWK * 32;
static constexpr int BTM = 64;
static constexpr int BNDMAX = 16;              // staged distinct destination rows per tile (more: read from global)

template <int NS, int NV>
struct BwdDims {
    static constexpr int MZ = NS + NV, DPRE = NS + 4 * NV, D = NS + 3 * NV;
    static constexpr int KZ = (MZ + 7) & ~7, KV = (NV + 7) & ~7, K1T = KZ + KV;      // T1 = [HZ | HG]
    static constexpr int KQ1 = K1T / 4, KQ3 = KV / 4, KQZ = KZ / 4;
    static constexpr int NSP = (NS + 7) & ~7, NDP = (NV + 7) & ~7;                   // accumulator widths
    // accumulator columns: [GS1 | GD] is ONE product (the two B tiles share the A operand HZ: a single N = NSP + NDP tile)
    static constexpr int cS1 = 0, cD = NSP, cS2 = NSP + NDP, cT = 2 * NSP + NDP, ACC = 256;
    static constexpr int SB = NSP / 8, VB = NDP / 8;                                 // 8-column blocks
    static constexpr int SQ = NS / 4, RS4 = NS % 4, NP = NV / 2, NU = SQ + (RS4 ? 1 : 0) + NP;
    static constexpr int halfT1 = BTM * K1T * 4, halfT3 = BTM * KV * 4, HALFB = halfT1 + 3 * halfT3, TBYTES = 2 * HALFB;
    static_assert(NS % 2 == 0 && NV % 2 == 0, "even channel counts (8-byte accesses)");
    static_assert(cT + 3 * NDP <= ACC && SB + VB <= 8 && NU <= 16, "tile shape");
    static_assert((DPRE & 1) == 0, "8-byte row alignment");
};

struct FusedBwdArgs {
    long long rows;
    const int* dst;            // [E] ascending
    const float* y;            // [E, 4]
    const float* pre1;         // [E rounded up to 64 rows, DPRE]: whole 64-row blocks are copied by cp.async.bulk
    const float* pre2;         // [E rounded up to 64 rows, DPRE]
    const float* gagg;         // [n_dst, D]
    const float* wz2;          // [(NS + NV), MZ]
    const float* wv2;          // [(NS + NV), NV]
    const float* nz2;
    const float* nv2;
    float* gpre1;              // [E rounded up to 64 rows, DPRE]: whole tiles are written by cp.async.bulk
    float* gpre2;              // [E rounded up to 64 rows, DPRE] or NULL
    float cs, cg;
};

template <int NS, int NV>
struct BwdSmem {
    using F = BwdDims<NS, NV>;
    static constexpr int o_bs1 = 0;                                         // [NSP + NDP][KZ]  hi | lo   (BZs rows, then BZd rows)
    static constexpr int o_bs2 = o_bs1 + 2 * (F::NSP + F::NDP) * F::KZ * 4; // [NSP][KV]
    static constexpr int o_bt = o_bs2 + 2 * F::NSP * F::KV * 4;             // [NDP][KV]
    static constexpr int o_t = (o_bt + 2 * F::NDP * F::KV * 4 + 1023) & ~1023;
    static constexpr int o_out = o_t + F::TBYTES;           // ONE operand set (see msg_fused_fwd.cu)
    static constexpr int TILEB = (BTM * F::DPRE * 4 + 127) & ~127;
    static constexpr int o_p2 = (o_out + TILEB + 127) & ~127;              // staged pre-activation of message 2, two tiles
    static constexpr int o_p1 = o_p2 + 2 * TILEB;                          // staged pre-activation of message 1, two tiles
    static constexpr int GAB = BNDMAX * F::D * 4;
    static constexpr int o_ga = o_p1 + 2 * TILEB;                          // staged cotangent rows of the distinct destinations, two tiles
    static constexpr int o_slot = o_ga + 2 * GAB;                          // row -> slot, two tiles
    static constexpr int o_g2t = (o_slot + 2 * (BTM + 4) * 4 + 127) & ~127;        // cotangent of message 2's pre-activation (bulk-stored)
    static constexpr int o_bar = o_g2t + TILEB;
    static constexpr int total = o_bar + 14 * 8 + 16;
};

__device__ __forceinline__ void bbulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bmbar_arrive_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bbulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bbulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bbulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bbulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int NS, int NV>
__global__ void __launch_bounds__(B_THREADS, 1) msg_fused_bwd_kernel(const __grid_constant__ FusedBwdArgs A) {
    using F = BwdDims<NS, NV>;
    using SM = BwdSmem<NS, NV>;
    constexpr int MZ = F::MZ, KZ = F::KZ, KV = F::KV, KQ1 = F::KQ1, KQ3 = F::KQ3, KQZ = F::KQZ, NSP = F::NSP, NDP = F::NDP;
    extern __shared__ __align__(1024) unsigned char smraw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + SM::o_bar);
    const uint32_t bar0 = smem_u32(bars);
    // barriers: 0,1 operand set full | 2,3 accumulator full | 4,5 accumulator empty | 6,10 staged rows of the build landed
    //           (even / odd tiles: the staging is double buffered, the copies are issued TWO tiles ahead) |
    //           7,8 staged pre-activation of message 1 landed (even / odd tiles) | 9 result tile stored (free) |
    //           11 g_pre2 tile stored (free)
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(BAR(i), BWK);
            mbar_init(BAR(2 + i), 1);
            mbar_init(BAR(4 + i), BWK);
        }
        mbar_init(BAR(6), 1);
        mbar_init(BAR(10), 1);
        mbar_init(BAR(11), 1);
        mbar_init(BAR(7), 1);
        mbar_init(BAR(8), 1);
        mbar_init(BAR(9), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        float4* z = reinterpret_cast<float4*>(smraw + SM::o_t);
        for (int t = tid; t < F::TBYTES >> 4; t += B_THREADS) z[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    {   // transposed weights -> canonical K-major B tiles (hi | lo): row n = input channel of message 2, K = its output
        // rows [n0, n0 + NN) of a tile of NT rows
        auto fill = [&](int off, int NT, int n0, int NN, int KK, int nvalid, int kvalid, auto value) {
            unsigned char* b = smraw + off;
            for (int t = tid; t < NN * KK; t += B_THREADS) {
                const int n = t / KK, k = t - n * KK;
                const float x = (n < nvalid && k < kvalid) ? value(n, k) : 0.0f;
                float hi, lo;
                split_tf32(x, hi, lo);
                const int o = canon_off(n0 + n, k, KK >> 2);
                *reinterpret_cast<float*>(b + o) = hi;
                *reinterpret_cast<float*>(b + NT * KK * 4 + o) = lo;
            }
        };
        auto nz = [&](int m) { return A.nz2 ? __ldg(A.nz2 + m) : 1.0f; };
        auto nv = [&](int m) { return A.nv2 ? __ldg(A.nv2 + 3 * m) : 1.0f; };
        fill(SM::o_bs1, NSP + NDP, 0, NSP, KZ, NS, MZ, [&](int k, int m) { return nz(m) * __ldg(A.wz2 + k * MZ + m); });
        fill(SM::o_bs1, NSP + NDP, NSP, NDP, KZ, NV, MZ, [&](int k, int m) { return nz(m) * C3f * __ldg(A.wz2 + (NS + k) * MZ + m); });
        fill(SM::o_bs2, NSP, 0, NSP, KV, NS, NV, [&](int k, int m) { return nv(m) * C3f * __ldg(A.wv2 + k * NV + m); });
        fill(SM::o_bt, NDP, 0, NDP, KV, NV, NV, [&](int k, int m) { return nv(m) * C3f * __ldg(A.wv2 + (NS + k) * NV + m); });
    }
    fence_proxy_async();
    if (warp == BWK) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const long long R = A.rows;
    const long long ntiles = (R + BTM - 1) / BTM;
    const int nt = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

    if (warp == BWK) {
        // ================= MMA issuer
        const uint32_t sb = smem_u32(smraw);
        const uint32_t idS = make_idesc(NSP), idD = make_idesc(NDP), idSD = make_idesc(NSP + NDP);
        constexpr uint32_t sboT1 = KQ1 * 128, sboT3 = KQ3 * 128, sboZ = KQZ * 128, sboG = KQ3 * 128;
        const uint64_t bS1h = make_desc(sb + SM::o_bs1, sboZ), bS1l = make_desc(sb + SM::o_bs1 + (NSP + NDP) * KZ * 4, sboZ);
        const uint64_t bS2h = make_desc(sb + SM::o_bs2, sboG), bS2l = make_desc(sb + SM::o_bs2 + NSP * KV * 4, sboG);
        const uint64_t bTh = make_desc(sb + SM::o_bt, sboT3), bTl = make_desc(sb + SM::o_bt + NDP * KV * 4, sboT3);
        constexpr uint64_t v3 = (uint64_t)(F::halfT3 >> 4);
        // TMA staging (this otherwise idle warp, a tile ahead): the tile's pre-activation blocks are contiguous in HBM (ONE cp.async.bulk each),
        // the cotangent rows are gathered once per distinct destination of the tile (dst is ascending)
        int pf_dst0 = 0, pf_dst1 = 0;
        int* sslot = reinterpret_cast<int*>(smraw + SM::o_slot);
        const uint32_t sm_u32 = smem_u32(smraw);
        auto load_pf = [&](int it) {
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * BTM;
            long long g0 = row0 + lane, g1 = g0 + 32;
            if (g0 > R - 1) g0 = R - 1;
            if (g1 > R - 1) g1 = R - 1;
            pf_dst0 = ldgi_v(A.dst + g0);
            pf_dst1 = ldgi_v(A.dst + g1);
        };
        auto issue_pf = [&](int it) {
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * BTM;
            const int up0 = __shfl_up_sync(0xffffffffu, pf_dst0, 1), last0 = __shfl_sync(0xffffffffu, pf_dst0, 31);
            const int up1 = __shfl_up_sync(0xffffffffu, pf_dst1, 1);
            const bool f0 = lane > 0 && pf_dst0 != up0;
            const bool f1 = pf_dst1 != (lane == 0 ? last0 : up1);
            const unsigned b0 = __ballot_sync(0xffffffffu, f0), b1 = __ballot_sync(0xffffffffu, f1);
            const unsigned le = 0xffffffffu >> (31 - lane);
            const int s0 = __popc(b0 & le), s1 = __popc(b0) + __popc(b1 & le);
            const int pb = it & 1;
            const uint32_t pbar = pb ? BAR(10) : BAR(6);
            sslot[pb * (BTM + 4) + lane] = s0;
            sslot[pb * (BTM + 4) + lane + 32] = s1;
            const int ndist = __popc(b0) + __popc(b1) + 1;
            const int ncopy = min(ndist, BNDMAX);
            if (lane == 0) sslot[pb * (BTM + 4) + BTM] = ndist > BNDMAX ? 1 : 0;
            constexpr uint32_t blk = BTM * F::DPRE * 4;
            if (lane == 0) {
                bmbar_arrive_tx(pbar, blk + ncopy * F::D * 4);
                bbulk_g2s(sm_u32 + SM::o_p2 + pb * SM::TILEB, A.pre2 + row0 * F::DPRE, blk, pbar);
            }
            __syncwarp();
            if ((lane == 0 || f0) && s0 < BNDMAX) bbulk_g2s(sm_u32 + SM::o_ga + pb * SM::GAB + s0 * F::D * 4, A.gagg + (long long)pf_dst0 * F::D, F::D * 4, pbar);
            if (f1 && s1 < BNDMAX) bbulk_g2s(sm_u32 + SM::o_ga + pb * SM::GAB + s1 * F::D * 4, A.gagg + (long long)pf_dst1 * F::D, F::D * 4, pbar);
        };
        // pre-activation of message 1 of tile `it` -> buffer it & 1 (read by drain(it), two iterations after this is issued:
        // the buffer is free once drain(it - 2) is complete, i.e. after the named barrier that follows it)
        auto issue_p1 = [&](int it) {
            if (lane == 0) {
                const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * BTM;
                constexpr uint32_t blk = BTM * F::DPRE * 4;
                bmbar_arrive_tx(BAR(7 + (it & 1)), blk);
                bbulk_g2s(sm_u32 + SM::o_p1 + (it & 1) * SM::TILEB, A.pre1 + row0 * F::DPRE, blk, BAR(7 + (it & 1)));
            }
        };
        if (nt > 0) {
            load_pf(0);
            issue_pf(0);
            if (nt > 1) { load_pf(1); issue_pf(1); }
            if (nt > 2) load_pf(2);
        }
        for (int it = 0; it < nt; ++it) {
            const int b = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            mbar_wait(BAR(b), ph);
            if (lane == 0 && A.gpre2) {        // cotangent of message 2's pre-activation of this tile: one bulk store; the prefetch
                const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * BTM;   // below waits for its reads
                bbulk_s2g(A.gpre2 + row0 * F::DPRE, sb + SM::o_g2t, BTM * F::DPRE * 4);
                bbulk_commit();
                bbulk_wait_read0();
                mbar_arrive(BAR(11));          // the next build may overwrite the tile
            }
            __syncwarp();
            if (it + 2 < nt) {                 // staging buffer it & 1 is free: every worker arrived from build(it).  Issued BEFORE
                issue_pf(it + 2);              // the MMAs (their issue takes ~3k cycles of this thread), TWO tiles ahead: a bulk
                if (it + 3 < nt) load_pf(it + 3);   // copy needs longer to land than the drain of one tile takes
            }
            mbar_wait(BAR(4 + b), ph ^ 1);     // drain(it - 2) done: accumulator b AND staging buffer b of message 1 are free
            issue_p1(it);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t acc = tmem_base + (uint32_t)b * F::ACC;
                const uint32_t t1 = sb + SM::o_t;
                const uint64_t a1h = make_desc(t1, sboT1), a1l = make_desc(t1 + F::HALFB, sboT1);
#pragma unroll
                for (int j = 0; j < KZ / 8; ++j) {
                    const uint64_t o = (uint64_t)(j * 16);
                    tc_mma_tf32(acc + F::cS1, a1h + o, bS1h + o, idSD, j ? 1u : 0u);     // [GS1 | GD]
                    tc_mma_tf32(acc + F::cS1, a1h + o, bS1l + o, idSD, 1u);
                    tc_mma_tf32(acc + F::cS1, a1l + o, bS1h + o, idSD, 1u);
                }
#pragma unroll
                for (int j = 0; j < KV / 8; ++j) {
                    const uint64_t o = (uint64_t)(j * 16), oa = (uint64_t)((KZ / 8 + j) * 16);
                    tc_mma_tf32(acc + F::cS2, a1h + oa, bS2h + o, idS, j ? 1u : 0u);
                    tc_mma_tf32(acc + F::cS2, a1h + oa, bS2l + o, idS, 1u);
                    tc_mma_tf32(acc + F::cS2, a1l + oa, bS2h + o, idS, 1u);
                }
                const uint64_t a3h = make_desc(t1 + F::halfT1, sboT3), a3l = make_desc(t1 + F::halfT1 + F::HALFB, sboT3);
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int j = 0; j < KV / 8; ++j) {
                        const uint64_t o = (uint64_t)(j * 16), oa = o + (uint64_t)c * v3;
                        tc_mma_tf32(acc + F::cT + c * NDP, a3h + oa, bTh + o, idD, j ? 1u : 0u);
                        tc_mma_tf32(acc + F::cT + c * NDP, a3h + oa, bTl + o, idD, 1u);
                        tc_mma_tf32(acc + F::cT + c * NDP, a3l + oa, bTh + o, idD, 1u);
                    }
                tc_commit(BAR(2 + b));
            }
            __syncwarp();
            if (it >= 1) {                     // result tile (g_pre1) of tile it-1: stored as soon as its drain is complete
                mbar_wait(BAR(4 + ((it - 1) & 1)), (uint32_t)(((it - 1) >> 1) & 1));
                if (lane == 0) {
                    const long long row0 = ((long long)blockIdx.x + (long long)(it - 1) * gridDim.x) * BTM;
                    bbulk_s2g(A.gpre1 + row0 * F::DPRE, sb + SM::o_out, BTM * F::DPRE * 4);
                    bbulk_commit();
                    bbulk_wait_read0();
                    mbar_arrive(BAR(9));
                }
                __syncwarp();
            }
        }
        if (nt > 0) {
            mbar_wait(BAR(4 + ((nt - 1) & 1)), (uint32_t)(((nt - 1) >> 1) & 1));
            if (lane == 0) {
                const long long row0 = ((long long)blockIdx.x + (long long)(nt - 1) * gridDim.x) * BTM;
                bbulk_s2g(A.gpre1 + row0 * F::DPRE, sb + SM::o_out, BTM * F::DPRE * 4);
                bbulk_commit();
                bbulk_wait0();
            }
            __syncwarp();
        }
    } else {
        // ================= workers
        const int* sslot = reinterpret_cast<const int*>(smraw + SM::o_slot);
        const int r8 = lane & 7, cq = lane >> 3;
        const int wrow = (warp & 7) * 8 + r8;
        const int sub = warp >> 3;
        const int rp1 = (((wrow >> 3) * KQ1) << 7) + ((wrow & 7) << 4);
        const int rp3 = (((wrow >> 3) * KQ3) << 7) + ((wrow & 7) << 4);
        int n_dst = 0;
        float4 n_y = make_float4(0.f, 0.f, 0.f, 0.f);
        auto load_row = [&](int it) {
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * BTM;
            long long gr = row0 + wrow;
            if (gr > R - 1) gr = R - 1;
            n_dst = ldgi_v(A.dst + gr);
            n_y = ldg4_v(A.y + 4 * gr);
        };
        auto st_hl4 = [&](unsigned char* p, float a, float b, float c, float d) {
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
        auto swish_vjp = [&](float g, float x) { const float s = sigm(x); return A.cs * g * s * fmaf(x, 1.0f - s, 1.0f); };
        // OVF: the tile has more than BNDMAX distinct destinations (rare): the cotangent rows past the staged slots come from
        // global memory through a generic pointer; in the common case every load of the build is an LDS
        auto build_body = [&](int it, auto OVF) {
            unsigned char* tset = smraw + SM::o_t;
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * BTM;
            long long gr = row0 + wrow;
            const bool valid = gr < R;
            if (!valid) gr = R - 1;
            const float4 y = n_y;
            const int pb = it & 1;
            const float* pre = reinterpret_cast<const float*>(smraw + SM::o_p2 + pb * SM::TILEB) + wrow * F::DPRE;
            const int slot = sslot[pb * (BTM + 4) + wrow];
            const float* gms = reinterpret_cast<const float*>(smraw + SM::o_ga + pb * SM::GAB) + (slot < BNDMAX ? slot : 0) * F::D;
            const float* gm = (decltype(OVF)::value && slot >= BNDMAX) ? A.gagg + (long long)n_dst * F::D : gms;
            float* go = reinterpret_cast<float*>(smraw + SM::o_g2t) + wrow * F::DPRE;   // tile rows, bulk-stored by the MMA warp
            const bool wr = A.gpre2 != nullptr;
            auto ld2s = [&](const float* q) { return *reinterpret_cast<const float2*>(q); };
            auto ld4s = [&](const float* q) { return *reinterpret_cast<const float4*>(q); };
#pragma unroll
            for (int round = 0; round < 2; ++round) {
                const int u = 8 * sub + 4 * round + cq;
                if (u >= F::NU) continue;
                if (u < F::SQ) {
                    const float2 x01 = ld2s(pre + 4 * u), x23 = ld2s(pre + 4 * u + 2);
                    const float4 g = ld4s(gm + 4 * u);
                    const float h0 = swish_vjp(g.x, x01.x), h1 = swish_vjp(g.y, x01.y), h2 = swish_vjp(g.z, x23.x), h3 = swish_vjp(g.w, x23.y);
                    if (wr) {
                        *reinterpret_cast<float2*>(go + 4 * u) = make_float2(h0, h1);
                        *reinterpret_cast<float2*>(go + 4 * u + 2) = make_float2(h2, h3);
                    }
                    st_hl4(tset + rp1 + (u << 7), h0, h1, h2, h3);
                } else if (F::RS4 && u == F::SQ) {
                    const float2 x = ld2s(pre + 4 * u), g = ld2s(gm + 4 * u);
                    const float h0 = swish_vjp(g.x, x.x), h1 = swish_vjp(g.y, x.y);
                    if (wr) *reinterpret_cast<float2*>(go + 4 * u) = make_float2(h0, h1);
                    st_hl2(tset + rp1 + (u << 7), h0, h1);
                } else {
                    const int i = u - F::SQ - (F::RS4 ? 1 : 0);
                    const int kg = NS + 2 * i;                       // gate channels kg, kg + 1 (vectors 2i, 2i + 1)
                    const float2 xg = ld2s(pre + kg);
                    const float* pv = pre + MZ + 6 * i;
                    const float* gv = gm + NS + 6 * i;
                    const float2 p01 = ld2s(pv), p23 = ld2s(pv + 2), p45 = ld2s(pv + 4);
                    const float2 g01 = ld2s(gv), g23 = ld2s(gv + 2), g45 = ld2s(gv + 4);
                    const float s0 = sigm(xg.x), s1 = sigm(xg.y);
                    const float d0 = fmaf(g01.x, p01.x, fmaf(g01.y, p01.y, g23.x * p23.x));
                    const float d1 = fmaf(g23.y, p23.y, fmaf(g45.x, p45.x, g45.y * p45.y));
                    const float hg0 = A.cg * s0 * (1.0f - s0) * d0, hg1 = A.cg * s1 * (1.0f - s1) * d1;
                    const float a0 = A.cg * s0, a1 = A.cg * s1;
                    const float q00 = a0 * g01.x, q01 = a0 * g01.y, q02 = a0 * g23.x, q10 = a1 * g23.y, q11 = a1 * g45.x, q12 = a1 * g45.y;
                    if (wr) {
                        *reinterpret_cast<float2*>(go + kg) = make_float2(hg0, hg1);
                        float* qo = go + MZ + 6 * i;
                        *reinterpret_cast<float2*>(qo) = make_float2(q00, q01);
                        *reinterpret_cast<float2*>(qo + 2) = make_float2(q02, q10);
                        *reinterpret_cast<float2*>(qo + 4) = make_float2(q11, q12);
                    }
                    // HZ gates at K slots kg, kg + 1; HG and HVc at K slots 2i, 2i + 1 of their tiles
                    st_hl2(tset + rp1 + ((kg >> 2) << 7) + ((kg & 3) << 2), hg0, hg1);
                    const int vo = ((i >> 1) << 7) + ((i & 1) << 3);
                    st_hl2(tset + rp1 + (KQZ << 7) + vo, fmaf(y.y, q00, fmaf(y.z, q01, y.w * q02)), fmaf(y.y, q10, fmaf(y.z, q11, y.w * q12)));
                    unsigned char* t3 = tset + F::halfT1 + rp3 + vo;
                    st_hl2(t3, q00, q10);
                    st_hl2(t3 + F::halfT3, q01, q11);
                    st_hl2(t3 + 2 * F::halfT3, q02, q12);
                }
            }
        };
        auto build = [&](int it) {
            const int pb = it & 1;
            mbar_wait(pb ? BAR(10) : BAR(6), (uint32_t)((it >> 1) & 1));
            if (it >= 1 && A.gpre2) mbar_wait(BAR(11), (uint32_t)((it - 1) & 1));    // the g_pre2 tile of tile it-1 has been stored
            if (it >= 1) mbar_wait(BAR(2 + ((it - 1) & 1)), (uint32_t)(((it - 1) >> 1) & 1));   // MMAs of tile it-1 have read the set
            if (sslot[pb * (BTM + 4) + BTM] == 0) build_body(it, std::false_type{});
            else build_body(it, std::true_type{});
        };
        // ---- epilogue: warp = (lane quarter e, task pair jq): tasks jq and jq + 4 of [SB scalar blocks | VB vector blocks]
        const int e = warp & 3, jq = warp >> 2;
        float* otile = reinterpret_cast<float*>(smraw + SM::o_out);
        const int fg = lane >> 2, fq = lane & 3;
        float4 ypre = make_float4(0.f, 0.f, 0.f, 0.f), ypre2 = ypre;
        auto task_kind = [&](int h, int& blk) { blk = jq + 4 * h; return blk < F::SB ? 0 : (blk < F::SB + F::VB ? 1 : 2); };
        auto prefetch_x = [&](int it) {   // SH rows of this thread's two epilogue rows (a = 16 e + fg, b = a + 8) of tile `it`
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * BTM;
            long long ga = row0 + 16 * e + fg, gb = ga + 8;
            if (ga > R - 1) ga = R - 1;
            if (gb > R - 1) gb = R - 1;
            ypre = ldg4_v(A.y + ga * 4);
            ypre2 = ldg4_v(A.y + gb * 4);
        };
        auto vec_out = [&](float* orow, int k, const float4& y, const float2& xg, const float2 (&xv)[3], float d0, float tx0,
                           float ty0, float tz0, float d1, float tx1, float ty1, float tz1) {
            // cotangent of the gated vectors k, k + 1 -> gate VJP of message 1
            const float v00 = fmaf(y.y, d0, y.x * tx0), v01 = fmaf(y.z, d0, y.x * ty0), v02 = fmaf(y.w, d0, y.x * tz0);
            const float v10 = fmaf(y.y, d1, y.x * tx1), v11 = fmaf(y.z, d1, y.x * ty1), v12 = fmaf(y.w, d1, y.x * tz1);
            const float s0 = sigm(xg.x), s1 = sigm(xg.y);
            // xv = (p00 p01 | p02 p10 | p11 p12)
            const float dot0 = fmaf(v00, xv[0].x, fmaf(v01, xv[0].y, v02 * xv[1].x));
            const float dot1 = fmaf(v10, xv[1].y, fmaf(v11, xv[2].x, v12 * xv[2].y));
            *reinterpret_cast<float2*>(orow + NS + k) = make_float2(A.cg * s0 * (1.0f - s0) * dot0, A.cg * s1 * (1.0f - s1) * dot1);
            const float a0 = A.cg * s0, a1 = A.cg * s1;
            float* ov = orow + MZ + 3 * k;
            *reinterpret_cast<float2*>(ov) = make_float2(a0 * v00, a0 * v01);
            *reinterpret_cast<float2*>(ov + 2) = make_float2(a0 * v02, a1 * v10);
            *reinterpret_cast<float2*>(ov + 4) = make_float2(a1 * v11, a1 * v12);
        };
        auto drain = [&](int it) {
            const int b = it & 1;
            const float4 ya = ypre, yb = ypre2;
            mbar_wait(BAR(2 + b), (it >> 1) & 1);
            mbar_wait(BAR(7 + (it & 1)), (uint32_t)((it >> 1) & 1));
            if (it >= 1) mbar_wait(BAR(9), (uint32_t)((it - 1) & 1));     // the result tile of tile it-1 has been stored
            tc_fence_after();
            const uint32_t acc = tmem_base + (uint32_t)b * F::ACC + ((uint32_t)(32 * e) << 16);
            float* oa = otile + (16 * e + fg) * F::DPRE;
            float* ob = oa + 8 * F::DPRE;
            // staged pre-activation of message 1 of this tile (it landed before the tile was built)
            const float* pa = reinterpret_cast<const float*>(smraw + SM::o_p1 + (it & 1) * SM::TILEB) + (16 * e + fg) * F::DPRE;
            const float* pb = pa + 8 * F::DPRE;
            auto l2 = [&](const float* q) { return *reinterpret_cast<const float2*>(q); };
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int blk;
                const int kind = task_kind(h, blk);
                if (kind == 0) {
                    float s1[4], s2[4];
                    tc_ld_16x256(acc + F::cS1 + 8 * blk, s1);
                    tc_ld_16x256(acc + F::cS2 + 8 * blk, s2);
                    tc_wait_ld();
                    const int k = 8 * blk + 2 * fq;
                    if (k < NS) {
                        const float2 xa = l2(pa + k), xb = l2(pb + k);
                        *reinterpret_cast<float2*>(oa + k) = make_float2(swish_vjp(fmaf(ya.x, s1[0], s2[0]), xa.x),
                                                                         swish_vjp(fmaf(ya.x, s1[1], s2[1]), xa.y));
                        *reinterpret_cast<float2*>(ob + k) = make_float2(swish_vjp(fmaf(yb.x, s1[2], s2[2]), xb.x),
                                                                         swish_vjp(fmaf(yb.x, s1[3], s2[3]), xb.y));
                    }
                } else if (kind == 1) {
                    float d[4], tx[4], ty[4], tz[4];
                    const int cb = 8 * (blk - F::SB);
                    tc_ld_16x256(acc + F::cD + cb, d);
                    tc_ld_16x256(acc + F::cT + cb, tx);
                    tc_ld_16x256(acc + F::cT + NDP + cb, ty);
                    tc_ld_16x256(acc + F::cT + 2 * NDP + cb, tz);
                    tc_wait_ld();
                    const int k = cb + 2 * fq;
                    if (k < NV) {
                        const float2 xva[3] = {l2(pa + MZ + 3 * k), l2(pa + MZ + 3 * k + 2), l2(pa + MZ + 3 * k + 4)};
                        const float2 xvb[3] = {l2(pb + MZ + 3 * k), l2(pb + MZ + 3 * k + 2), l2(pb + MZ + 3 * k + 4)};
                        vec_out(oa, k, ya, l2(pa + NS + k), xva, d[0], tx[0], ty[0], tz[0], d[1], tx[1], ty[1], tz[1]);
                        vec_out(ob, k, yb, l2(pb + NS + k), xvb, d[2], tx[2], ty[2], tz[2], d[3], tx[3], ty[3], tz[3]);
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async();               // the result tile is read by a bulk store (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(4 + b));
        };
        if (nt > 0) load_row(0);
        for (int it = 0; it < nt; ++it) {
            build(it);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(it & 1));
            if (it + 1 < nt) load_row(it + 1);
            // no CTA-wide barrier: every hand-off is an mbarrier (staged rows and slots: issued by the MMA warp once all
            // workers arrived from the build; operand set: MMAs of the previous tile; result tile: its bulk store)
            if (it >= 1) drain(it - 1);
            prefetch_x(it);
        }
        if (nt > 0) drain(nt - 1);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == BWK) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

template <int NS, int NV>
static int msg_fused_bwd_launch(const FusedBwdArgs& A, cudaStream_t st) {
    using SM = BwdSmem<NS, NV>;
    int dev = 0, maxsm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (SM::total > maxsm) { set_error("msg_fused_backward: %d bytes of shared memory needed, %d available", SM::total, maxsm); return SE3_ERR_TOO_LARGE; }
    const int smem = std::max(SM::total, 120 * 1024);
    static bool attr_set = false;
    if (!attr_set) {
        SE3_CUDA_TRY(cudaFuncSetAttribute(msg_fused_bwd_kernel<NS, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
        attr_set = true;
    }
    const long long ntiles = (A.rows + BTM - 1) / BTM;
    const int grid = (int)std::min<long long>(ntiles, num_sms());
    msg_fused_bwd_kernel<NS, NV><<<grid, B_THREADS, smem, st>>>(A);
    SE3_LAUNCHED();
    g_tc_launches.fetch_add(1, std::memory_order_relaxed);
    return SE3_OK;
}

}  // namespace se3

using namespace se3;

extern "C" int se3_msg_fused_backward(int32_t ns, int32_t nv, int64_t rows, const int32_t* dst, const float* y,
                                      const float* pre1, const float* pre2, const float* gagg, const float* wz2,
                                      const float* wv2, const float* nz2, const float* nv2, float gate_cs, float gate_cg,
                                      float* gpre1, float* gpre2, void* stream) {
    if (rows < 0 || rows >= (1ll << 31) - BTM) { set_error("msg_fused_backward: bad row count"); return SE3_ERR_INVALID; }
    if (rows == 0) return SE3_OK;
    if (!dst || !y || !pre1 || !pre2 || !gagg || !wz2 || !wv2 || !gpre1) { set_error("msg_fused_backward: null argument"); return SE3_ERR_INVALID; }
    if (((uintptr_t)y | (uintptr_t)gagg | (uintptr_t)gpre1) & 15) { set_error("msg_fused_backward: 16-byte alignment"); return SE3_ERR_INVALID; }
    FusedBwdArgs A;
    A.rows = rows; A.dst = dst; A.y = y; A.pre1 = pre1; A.pre2 = pre2; A.gagg = gagg; A.wz2 = wz2; A.wv2 = wv2; A.nz2 = nz2;
    A.nv2 = nv2; A.gpre1 = gpre1; A.gpre2 = gpre2; A.cs = gate_cs; A.cg = gate_cg;
    if (ns == 34 && nv == 10) return msg_fused_bwd_launch<34, 10>(A, (cudaStream_t)stream);
    if (ns == 16 && nv == 8) return msg_fused_bwd_launch<16, 8>(A, (cudaStream_t)stream);
    set_error("msg_fused_backward: hidden irreps %dx0e+%dx1o are not instantiated", (int)ns, (int)nv);
    return SE3_ERR_INVALID;
}
