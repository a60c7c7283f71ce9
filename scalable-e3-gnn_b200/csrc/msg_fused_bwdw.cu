// Weight-gradient kernel of message 2 of the SEGNN message layer on sm_100a (tcgen05, 3xTF32, accumulators resident in
// TMEM for all tiles of a CTA).  Autograd of L1TensorProduct.forward (L1TP:242-297) with respect to weights_l0e /
// weights_l1o, for in1 = the gated message 1 (saved) and the cotangent g_pre2 that csrc/msg_fused_bwd.cu writes.
//
// The reduction runs over the edges, so rows are the MMA K dimension and the row-major tiles are MN-major operands in
// the layout tools/probe/mma_probe2.cu established (SWIZZLE_128B_BASE32B: 128-byte column chunks of 32 slots, 4-row
// atoms, 32-byte units XOR-ed with row & 3, chunk stride = LBO; see l1tp_tc2_bwdw.cu).  Per row, with S / V_kc the
// scalars / vector components of message 1 and g0 / g1 the scalar / vector part of g_pre2 (norms and 1/sqrt(3) are
// applied once, in the final reduction):
//     M side:  T1 = [Y0 g0 | sum_c Y1c g1[.][c]]   HZ = g0           HVc = Y0 g1[.][c]
//     N side:  S                                   D  = <V_k, Y1>    Vc
//     A1 [T1 x S]  += T1^T S      A2 [HZ x D] += HZ^T D      A3c [HVc x Vc] += HVc^T Vc
// issued as TWO products per 8 rows: [T1 | HZ]^T [S | D] (M = 128, N = 64: the two wanted blocks of a 2 x 2 block
// accumulator) and [HVx | HVy | HVz]^T [Vx | Vy | Vz] (M = 64, N = 48: the three diagonal blocks); the cross blocks are
// simply not read.  That is 6 tcgen05.mma per 8 rows instead of 15: the kernel is bound by the single issuing thread
// (the first version with one product per block took as long as the kernel it replaces, 0.76 ms per 1.78M rows).
// The output-channel side is the MMA M dimension and the input-channel side is N.
// Feeding: the tile's rows of message 1 and of g_pre2 are contiguous in HBM: TWO cp.async.bulk (TMA) copies per
// 32-row tile into a double-buffered staging area, issued a tile ahead; the 16 worker warps only rescale by the SH,
// split into tf32 hi / lo and store 16-byte pieces (lane = (piece, row of a 4-row atom): conflict-free).
#include <algorithm>

#include "tc_common.cuh"

namespace se3 {

static constexpr int WW = 16;                  // worker warps
static constexpr int W_THREADS = (WW + 1) * 32;
static constexpr int WWT = WW * 32;
static constexpr int WTW = 32;                 // rows per tile (4 MMA K-steps)
static constexpr int WCH = WTW * 128;          // bytes of one 32-slot chunk
// chunk ids inside an operand set (hi part; lo follows at +WHALF)
static constexpr int kT1 = 0, kHZ = 2, kHV = 4, kS = 6, kV = 8, WNCHK = 10;
static constexpr int WHALF = WNCHK * WCH, WSET = 2 * WHALF;

template <int NS, int NV>
struct WDims {
    static constexpr int MZ = NS + NV, DPRE = NS + 4 * NV, D = NS + 3 * NV;
    static constexpr int N1 = (NS + 7) & ~7;                       // S slots (D sits at slot 48 of the same N = 64 operand)
    static constexpr int cA12 = 0, cA3 = 64, NCOL = 64 + 48;       // TMEM columns: [T1 | HZ] x [S | D], then HV x V
    static constexpr int ROWS = 128 + 64;                          // rows of a partial: 128 lanes of A12, 64 slots of A3
    static constexpr int GP = (MZ + 3) / 4, HP = (NV + 3) / 4, SP = (NS + 3) / 4, VP = (NV + 3) / 4;   // 16-byte pieces
    static constexpr int PART = 128 * 64 + 64 * 48;                // floats of one CTA's partial: A12 [128][64] | A3 [64][48]
    static_assert(MZ + NV <= 64 && MZ % 4 == 0, "T1 = [Y0 g0 | HG] fits 64 slots, HG starts on a piece boundary");
    static_assert(NV <= 16 && NS + 8 <= 48 && N1 <= 40, "vector operands use 16 slots; D sits at slot 16 of S's second chunk");
    static_assert(NS % 2 == 0 && NV % 2 == 0 && DPRE % 2 == 0, "8-byte aligned rows");
    static_assert(GP + HP <= 16 && SP + VP <= 16, "two rounds of eight tasks per warp pair");
    static_assert(NCOL <= 128 && N1 <= 48, "TMEM columns; S and D share one N = 64 operand");
};

struct FusedBwdWArgs {
    long long rows;
    const float* y;            // [E, 4]
    const float* m1;           // [E rounded up to 32 rows, D]
    const float* gpre2;        // [E rounded up to 32 rows, DPRE]
    float* partials;           // [grid, 64, NCOL]
};

template <int NS, int NV>
struct WSmem {
    using F = WDims<NS, NV>;
    static constexpr int o_set = 0;                                   // 2 operand sets (1024-byte aligned)
    static constexpr int M1B = (WTW * F::D * 4 + 127) & ~127, G2B = (WTW * F::DPRE * 4 + 127) & ~127;
    static constexpr int o_m1 = 2 * WSET;                             // 2 staged message-1 blocks
    static constexpr int o_g2 = o_m1 + 2 * M1B;                       // 2 staged cotangent blocks
    static constexpr int o_bar = o_g2 + 2 * G2B;
    static constexpr int total = o_bar + 12 * 8 + 16;
};

__device__ __forceinline__ int wtw_off(int chunk, int row, int slot) {   // byte offset inside the hi part of a set
    return chunk * WCH + (slot >> 5) * WCH + row * 128 + (((((slot & 31) >> 3)) ^ (row & 3)) << 5) + ((slot & 7) << 2);
}
__device__ __forceinline__ uint64_t wmk_desc_mn(uint32_t saddr) {        // MN-major, SWIZZLE_128B_BASE32B
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((WCH >> 4) & 0x3FFF) << 16) | ((uint64_t)((512 >> 4) & 0x3FFF) << 32) |
           (1ull << 46) | (1ull << 61);
}
__device__ __forceinline__ void wbulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void wmbar_arrive_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(bar), "r"(bytes) : "memory");
}

template <int NS, int NV>
__global__ void __launch_bounds__(W_THREADS, 1) msg_fused_bwdw_kernel(const __grid_constant__ FusedBwdWArgs A) {
    using F = WDims<NS, NV>;
    using SM = WSmem<NS, NV>;
    constexpr int MZ = F::MZ;
    extern __shared__ __align__(1024) unsigned char smraw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + SM::o_bar);
    const uint32_t bar0 = smem_u32(bars);
    // barriers: 0,1 set full | 2,3 set empty (MMAs done) | 4,5 staged rows landed | 6 accumulators final
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(BAR(i), WW);
            mbar_init(BAR(2 + i), 1);
            mbar_init(BAR(4 + i), 1);
        }
        mbar_init(BAR(6), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // zero both operand sets once: unused slots are never written again and must stay finite (0 * x)
        float4* z = reinterpret_cast<float4*>(smraw + SM::o_set);
        for (int t = tid; t < (2 * WSET) >> 4; t += W_THREADS) z[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    fence_proxy_async();
    if (warp == WW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long long R = A.rows;
    const long long ntiles = (R + WTW - 1) / WTW;
    const int nt = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

    if (warp == WW) {
        // ================= MMA issuer
        const uint32_t sb = smem_u32(smraw) + SM::o_set;
        const uint32_t id1 = make_idesc_ex(128, 64, 1, 1), id2 = make_idesc_ex(64, 48, 1, 1);
        for (int it = 0; it < nt; ++it) {
            const int b = it & 1;
            mbar_wait(BAR(b), (it >> 1) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t s0 = sb + (uint32_t)b * WSET;
                // M = 128 operand [T1 | HZ] = chunks kT1 .. kT1 + 3; N = 64 operand [S | - | D | -] = chunks kS, kS + 1
                // M = 64 operand [HVx | HVy | HVz | -] = chunks kHV, kHV + 1; N = 48 operand [Vx | Vy | Vz] = chunks kV, kV + 1
                const uint64_t dT = wmk_desc_mn(s0 + kT1 * WCH), dS = wmk_desc_mn(s0 + kS * WCH);
                const uint64_t dHV = wmk_desc_mn(s0 + kHV * WCH), dV = wmk_desc_mn(s0 + kV * WCH);
                const uint64_t lo = (uint64_t)(WHALF >> 4);
#pragma unroll
                for (int ks = 0; ks < WTW / 8; ++ks) {
                    const uint32_t acc0 = (it == 0 && ks == 0) ? 0u : 1u;
                    const uint64_t ko = (uint64_t)(ks * 64);   // 8 rows x 128 B, in 16-byte units
                    tc_mma_tf32(tmem_base + F::cA12, dT + ko, dS + ko, id1, acc0);
                    tc_mma_tf32(tmem_base + F::cA12, dT + ko, dS + ko + lo, id1, 1u);
                    tc_mma_tf32(tmem_base + F::cA12, dT + ko + lo, dS + ko, id1, 1u);
                    tc_mma_tf32(tmem_base + F::cA3, dHV + ko, dV + ko, id2, acc0);
                    tc_mma_tf32(tmem_base + F::cA3, dHV + ko, dV + ko + lo, id2, 1u);
                    tc_mma_tf32(tmem_base + F::cA3, dHV + ko + lo, dV + ko, id2, 1u);
                }
                tc_commit(BAR(2 + b));
            }
            __syncwarp();
        }
        if (lane == 0) tc_commit(BAR(6));
        __syncwarp();
    } else {
        // ================= workers.  lane = (task slot pc = lane & 7, row of the 4-row atom r4 = lane >> 3); warp w: row quad
        // w & 7 (rows 4 (w & 7) + r4), half w >> 3: half 0 builds the M side (cotangent), half 1 the N side (message 1)
        const int pc = lane & 7, r4 = lane >> 3;
        const int row = (warp & 7) * 4 + r4, half = warp >> 3;
        const uint32_t sm_u32 = smem_u32(smraw);
        float4 n_y = make_float4(0.f, 0.f, 0.f, 0.f);
        auto load_y = [&](int it) {
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * WTW;
            long long gr = row0 + row;
            if (gr > R - 1) gr = R - 1;
            n_y = ldg4_v(A.y + 4 * gr);
        };
        auto issue_pf = [&](int it) {      // warp 0, lane 0: the two contiguous blocks of tile `it` -> staging buffer it & 1
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * WTW;
            const int b = it & 1;
            wmbar_arrive_tx(BAR(4 + b), WTW * (F::D + F::DPRE) * 4);
            wbulk_g2s(sm_u32 + SM::o_m1 + b * SM::M1B, A.m1 + row0 * F::D, WTW * F::D * 4, BAR(4 + b));
            wbulk_g2s(sm_u32 + SM::o_g2 + b * SM::G2B, A.gpre2 + row0 * F::DPRE, WTW * F::DPRE * 4, BAR(4 + b));
        };
        auto st4 = [&](unsigned char* set, int off, float a, float b, float c, float d) {
            float4 h, l;
            split_tf32(a, h.x, l.x); split_tf32(b, h.y, l.y); split_tf32(c, h.z, l.z); split_tf32(d, h.w, l.w);
            *reinterpret_cast<float4*>(set + off) = h;
            *reinterpret_cast<float4*>(set + off + WHALF) = l;
        };
        auto l2 = [&](const float* q) { return *reinterpret_cast<const float2*>(q); };
        auto build = [&](int it) {
            const int b = it & 1;
            unsigned char* set = smraw + SM::o_set + b * WSET;
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * WTW;
            const bool valid = row0 + row < R;     // rows past the end contribute exact zeros (the staged bytes there are arbitrary)
            const float4 y = n_y;
            const float* g2 = reinterpret_cast<const float*>(smraw + SM::o_g2 + b * SM::G2B) + row * F::DPRE;
            const float* m1 = reinterpret_cast<const float*>(smraw + SM::o_m1 + b * SM::M1B) + row * F::D;
#pragma unroll
            for (int round = 0; round < 2; ++round) {
                const int t = 8 * round + pc;
                if (half == 0) {
                    if (t < F::GP) {
                        // four scalar cotangents g0[4t..4t+3] -> HZ piece t, T1 piece t (x Y0)
                        float2 a = l2(g2 + 4 * t), c = l2(g2 + 4 * t + 2);
                        if (!valid) { a = make_float2(0.f, 0.f); c = a; }
                        st4(set, wtw_off(kHZ, row, 4 * t), a.x, a.y, c.x, c.y);
                        st4(set, wtw_off(kT1, row, 4 * t), y.x * a.x, y.x * a.y, y.x * c.x, y.x * c.y);
                    } else if (t < F::GP + F::HP) {
                        // vector cotangents g1[m][c], m = 4q..4q+3 -> HVc piece q (x Y0), HG piece (sum_c Y1c g1[m][c])
                        const int q = t - F::GP;
                        float g[12];
#pragma unroll
                        for (int i = 0; i < 6; ++i) {
                            const bool in = 4 * q + (2 * i) / 3 < NV;      // the last piece may hold fewer than 4 channels
                            const float2 v = in ? l2(g2 + MZ + 12 * q + 2 * i) : make_float2(0.f, 0.f);
                            g[2 * i] = valid ? v.x : 0.f; g[2 * i + 1] = valid ? v.y : 0.f;
                        }
                        st4(set, wtw_off(kHV, row, 4 * q), y.x * g[0], y.x * g[3], y.x * g[6], y.x * g[9]);
                        st4(set, wtw_off(kHV, row, 16 + 4 * q), y.x * g[1], y.x * g[4], y.x * g[7], y.x * g[10]);
                        st4(set, wtw_off(kHV + 1, row, 4 * q), y.x * g[2], y.x * g[5], y.x * g[8], y.x * g[11]);
                        st4(set, wtw_off(kT1, row, MZ + 4 * q),
                            fmaf(y.y, g[0], fmaf(y.z, g[1], y.w * g[2])), fmaf(y.y, g[3], fmaf(y.z, g[4], y.w * g[5])),
                            fmaf(y.y, g[6], fmaf(y.z, g[7], y.w * g[8])), fmaf(y.y, g[9], fmaf(y.z, g[10], y.w * g[11])));
                    }
                } else {
                    if (t < F::SP) {
                        // four scalars of message 1 -> S piece t
                        float4 s = *reinterpret_cast<const float4*>(m1 + 4 * t);
                        if (4 * t + 2 >= NS) { s.z = 0.f; s.w = 0.f; }      // the last piece holds NS % 4 (= 2) scalars
                        if (!valid) s = make_float4(0.f, 0.f, 0.f, 0.f);
                        st4(set, wtw_off(kS, row, 4 * t), s.x, s.y, s.z, s.w);
                    } else if (t < F::SP + F::VP) {
                        // vector channels k = 4q..4q+3 -> Vc piece q, D piece q (<V_k, Y1>)
                        const int q = t - F::SP;
                        float v[12];
#pragma unroll
                        for (int i = 0; i < 6; ++i) {
                            const bool in = 4 * q + (2 * i) / 3 < NV;
                            const float2 u = in ? l2(m1 + NS + 12 * q + 2 * i) : make_float2(0.f, 0.f);
                            v[2 * i] = valid ? u.x : 0.f; v[2 * i + 1] = valid ? u.y : 0.f;
                        }
                        st4(set, wtw_off(kV, row, 4 * q), v[0], v[3], v[6], v[9]);
                        st4(set, wtw_off(kV, row, 16 + 4 * q), v[1], v[4], v[7], v[10]);
                        st4(set, wtw_off(kV + 1, row, 4 * q), v[2], v[5], v[8], v[11]);
                        st4(set, wtw_off(kS + 1, row, 16 + 4 * q),
                            fmaf(y.y, v[0], fmaf(y.z, v[1], y.w * v[2])), fmaf(y.y, v[3], fmaf(y.z, v[4], y.w * v[5])),
                            fmaf(y.y, v[6], fmaf(y.z, v[7], y.w * v[8])), fmaf(y.y, v[9], fmaf(y.z, v[10], y.w * v[11])));
                    }
                }
            }
        };

        if (nt > 0) {
            load_y(0);
            if (warp == 0 && lane == 0) {
                issue_pf(0);
                if (nt > 1) issue_pf(1);
            }
        }
        for (int it = 0; it < nt; ++it) {
            const int b = it & 1;
            mbar_wait(BAR(4 + b), (it >> 1) & 1);                // staged rows of this tile landed
            mbar_wait(BAR(2 + b), ((it >> 1) & 1) ^ 1);          // the MMAs that read this set two tiles ago are done
            build(it);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(b));
            if (it + 1 < nt) load_y(it + 1);
            named_bar(1, WWT);                                   // every worker is done with staging buffer b
            if (warp == 0 && lane == 0 && it + 2 < nt) issue_pf(it + 2);
        }
        // ---------------- final epilogue (warps 0-3): TMEM accumulators -> this CTA's partial: A12 [128][64] | A3 [64][48]
        if (warp < 4) {
            mbar_wait(BAR(6), 0);
            tc_fence_after();
            float* part = A.partials + (long long)blockIdx.x * F::PART;
            const uint32_t tq = tmem_base + ((uint32_t)(32 * warp) << 16);
            {   // M = 128: TMEM lane = row of [T1 | HZ]
                float4* o = reinterpret_cast<float4*>(part + (32 * warp + lane) * 64);
                for (int c0 = 0; c0 < 64; c0 += 8) {
                    float a[8];
                    tc_ld8(tq + F::cA12 + c0, a);
                    tc_wait_ld();
                    o[c0 / 4] = make_float4(a[0], a[1], a[2], a[3]);
                    o[c0 / 4 + 1] = make_float4(a[4], a[5], a[6], a[7]);
                }
            }
            {   // M = 64: slot 16 q + i lives in TMEM lane 32 q + i
                float4* o = reinterpret_cast<float4*>(part + 128 * 64 + (16 * warp + (lane & 15)) * 48);
                for (int c0 = 0; c0 < 48; c0 += 8) {
                    float a[8];
                    tc_ld8(tq + F::cA3 + c0, a);
                    tc_wait_ld();
                    if (lane < 16) {
                        o[c0 / 4] = make_float4(a[0], a[1], a[2], a[3]);
                        o[c0 / 4 + 1] = make_float4(a[4], a[5], a[6], a[7]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WW) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base) : "memory");
    }
}

// gwz [(NS + NV), MZ], gwv [(NS + NV), NV] (overwritten): fixed-order sum of the per-CTA partials, norms and 1/sqrt(3)
template <int NS, int NV>
__global__ void __launch_bounds__(256) msg_fused_bwdw_reduce_kernel(const float* __restrict__ part, int nparts, const float* nz,
                                                                    const float* nvn, float* __restrict__ gwz, float* __restrict__ gwv) {
    using F = WDims<NS, NV>;
    constexpr int MZ = F::MZ, CH = MZ + NV, ROWS = NS + NV;
    // one warp per output element: lane l sums partials l, l + 32, ... (fixed order), then a shuffle tree
    const int lane = threadIdx.x & 31;
    for (int t = blockIdx.x * 8 + (threadIdx.x >> 5); t < ROWS * CH; t += gridDim.x * 8) {
        const int k = t / CH, ch = t - k * CH;          // k: input channel (scalars, then vectors); ch: output channel
        float g = 0.0f;
        if (k < NS) {
            const float* p = part + ch * 64 + k;                                   // A12[T1 slot ch][S slot k]
            for (int q = lane; q < nparts; q += 32) g += p[(long long)q * F::PART];
        } else if (ch < MZ) {
            const float* p = part + (64 + ch) * 64 + 48 + (k - NS);                // A12[HZ slot ch][D slot]
            for (int q = lane; q < nparts; q += 32) g += p[(long long)q * F::PART];
            g *= C3f;
        } else {
            for (int c = 0; c < 3; ++c) {                                          // A3[HVc slot][Vc slot]: diagonal block c
                const float* p = part + 128 * 64 + (16 * c + ch - MZ) * 48 + 16 * c + (k - NS);
                for (int q = lane; q < nparts; q += 32) g += p[(long long)q * F::PART];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
        if (lane == 0) {
            if (ch < MZ) gwz[k * MZ + ch] = g * (nz ? nz[ch] : 1.0f);
            else gwv[k * NV + ch - MZ] = g * C3f * (nvn ? nvn[3 * (ch - MZ)] : 1.0f);
        }
    }
}

template <int NS, int NV>
static int msg_fused_bwdw_launch(const FusedBwdWArgs& A0, const float* nz, const float* nvn, float* gwz, float* gwv,
                                 int max_parts, cudaStream_t st) {
    using SM = WSmem<NS, NV>;
    int dev = 0, maxsm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (SM::total > maxsm) { set_error("msg_fused_bwdw: %d bytes of shared memory needed, %d available", SM::total, maxsm); return SE3_ERR_TOO_LARGE; }
    static bool attr_set = false;
    if (!attr_set) {
        SE3_CUDA_TRY(cudaFuncSetAttribute(msg_fused_bwdw_kernel<NS, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
        attr_set = true;
    }
    const long long ntiles = (A0.rows + WTW - 1) / WTW;
    const int grid = (int)std::min<long long>(ntiles, std::min(num_sms(), max_parts));
    msg_fused_bwdw_kernel<NS, NV><<<grid, W_THREADS, std::max(SM::total, 120 * 1024), st>>>(A0);
    SE3_LAUNCHED();
    g_tc_launches.fetch_add(1, std::memory_order_relaxed);
    msg_fused_bwdw_reduce_kernel<NS, NV><<<num_sms(), 256, 0, st>>>(A0.partials, grid, nz, nvn, gwz, gwv);
    SE3_LAUNCHED();
    return SE3_OK;
}

}  // namespace se3

using namespace se3;

extern "C" int se3_msg_fused_bwdw_parts(int32_t ns, int32_t nv, int32_t* max_parts, int32_t* part_floats) {
    if (!max_parts || !part_floats) { set_error("msg_fused_bwdw_parts: null argument"); return SE3_ERR_INVALID; }
    *max_parts = num_sms();
    *part_floats = 128 * 64 + 64 * 48;
    return SE3_OK;
}

extern "C" int se3_msg_fused_backward_w(int32_t ns, int32_t nv, int64_t rows, const float* y, const float* m1,
                                        const float* gpre2, const float* nz2, const float* nv2, float* gwz2, float* gwv2,
                                        float* partials, int32_t max_parts, void* stream) {
    if (rows < 0 || rows >= (1ll << 31) - WTW || !gwz2 || !gwv2 || max_parts < 1) { set_error("msg_fused_backward_w: bad argument"); return SE3_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    if (rows == 0) {
        SE3_CUDA_TRY(cudaMemsetAsync(gwz2, 0, sizeof(float) * (ns + nv) * (ns + nv), st));
        SE3_CUDA_TRY(cudaMemsetAsync(gwv2, 0, sizeof(float) * (ns + nv) * nv, st));
        return SE3_OK;
    }
    if (!y || !m1 || !gpre2 || !partials) { set_error("msg_fused_backward_w: null argument"); return SE3_ERR_INVALID; }
    if (((uintptr_t)y | (uintptr_t)m1 | (uintptr_t)gpre2 | (uintptr_t)partials) & 15) { set_error("msg_fused_backward_w: 16-byte alignment"); return SE3_ERR_INVALID; }
    FusedBwdWArgs A;
    A.rows = rows; A.y = y; A.m1 = m1; A.gpre2 = gpre2; A.partials = partials;
    if (ns == 34 && nv == 10) return msg_fused_bwdw_launch<34, 10>(A, nz2, nv2, gwz2, gwv2, max_parts, st);
    if (ns == 16 && nv == 8) return msg_fused_bwdw_launch<16, 8>(A, nz2, nv2, gwz2, gwv2, max_parts, st);
    set_error("msg_fused_backward_w: hidden irreps %dx0e+%dx1o are not instantiated", (int)ns, (int)nv);
    return SE3_ERR_INVALID;
}
