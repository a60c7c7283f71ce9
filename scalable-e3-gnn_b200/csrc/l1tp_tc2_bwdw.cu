// Second-generation tensor-core (tcgen05 / TMEM, 3xTF32) WEIGHT-gradient kernel of the fused l<=1 tensor-product
// layer (SEGNN case).
//
// The reduction runs over rows, so rows are the MMA K dimension.  The first-generation kernel (l1tp_tc_bwd.cu) stages
// the rows with cp.async / TMA and TRANSPOSES them into K-major tiles with 20 builder warps, because tf32 operands in
// the plain (no-swizzle) MN-major layout are rejected by the hardware.  tools/probe/mma_probe2.cu established the
// layout that IS accepted: MN-major, layout type 1 (SWIZZLE_128B_BASE32B): row-major [k][mn] tiles made of 128-byte
// column chunks, 4-row atoms (SBO = 512 B), chunk stride LBO, the 32-byte units of a row XOR-swizzled with (row & 3);
// operand start addresses may sit at 32-byte offsets inside a chunk.  With it the tiles are plain ROW-major:
//   * homogeneous skeleton of l1tp_tc2.cu: 16 worker warps gather the rows of tile t+1 from global memory into
//     registers (8 lanes x 16 B = one 128-byte line per row, 4 rows per warp instruction), build the operands and
//     store them once with 8/16-byte vector stores; no staging copy, no producer warps, no transposition;
//   * operand sets are double buffered per 32-row tile; the MMA warp issues 60 MMAs per tile into accumulators that
//     stay resident in TMEM for ALL tiles of the CTA and are written once, as per-CTA partials.
// Per row, with H the cotangent of the pre-activation (gate VJP and norms folded in, as in l1tp_tc2_bwd.cu):
//   A side: S (scalars), D = c3 <v, Y1>, AVc = c3 Y0 v_c          B side: T1 = [Y0 HZ | HG], T2 = HZ, T3c = HVc
//   gWZ_s|gWV_s += S^T T1     gWZ_d += D^T T2     gWV_v += sum_c AVc^T T3c
#include <algorithm>
#include <vector>

#include "tc_common.cuh"

namespace se3 {

static constexpr int W3 = 16;
static constexpr int T3_THREADS = (W3 + 1) * 32;
static constexpr int TW = 32;                     // rows per tile (4 MMA K-steps)
static constexpr int CH = TW * 128;               // bytes of one 32-slot chunk
// chunk ids inside an operand set (hi part; the lo part follows at +HALF)
static constexpr int cS = 0, cD = 3, cAV = 4, cT1 = 7, cT2 = 9, cT3 = 11, NCHK = 13;
static constexpr int HALFB = NCHK * CH, SETB = 2 * HALFB;

struct SegW {
    const float* base;
    const int32_t* idx;
    int ld, nss, nvs, koff, kdoff, wide, w, vcol;   // vcol: first vector column; koff: S slot of scalar 0 (wide)
};
struct WarpT {                                     // the (at most one of each kind) tasks of a worker warp
    short xs_seg, xs_q, xv_seg, xv_q, hs_q, hx_q, ex_seg, pad;
};
struct Tc3Args {
    long long rows;
    const float* in2;
    const float* nz;
    const float* nv;
    EpiL epi;
    const float* raw;
    const float* gout;
    const int32_t* gout_idx;
    float* partials;
    int wtot, gw_z_off, gw_v_off;
    int ns, nd, mz, mv, d_out, gwidth, oz0, ov0, nsz;
    int K1, MP;                                    // S slots, M of the S MMA (64 / 128)
    int N2, N3;
    int o_set, o_norm, o_bar;
    SegW seg[SE3_MAX_SEG];
    int nseg;
    WarpT wt[W3];
    short sl2ch[96];
    short exslot[4];                               // S slots of the columns of the narrow segment
};

__device__ __forceinline__ int tw_off(int chunk, int row, int slot) {   // byte offset inside the hi part of a set
    return chunk * CH + (slot >> 5) * CH + row * 128 + (((((slot & 31) >> 3)) ^ (row & 3)) << 5) + ((slot & 7) << 2);
}
__device__ __forceinline__ uint64_t mk_desc_mn(uint32_t saddr) {        // MN-major, SWIZZLE_128B_BASE32B
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((CH >> 4) & 0x3FFF) << 16) | ((uint64_t)((512 >> 4) & 0x3FFF) << 32) |
           (1ull << 46) | (1ull << 61);
}

template <bool GATE>
__global__ void __launch_bounds__(T3_THREADS, 1) l1tp_tc2_bwdw_kernel(const __grid_constant__ Tc3Args A) {
    extern __shared__ __align__(1024) unsigned char smraw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* norm = reinterpret_cast<float*>(smraw + A.o_norm);     // nz[mz] then nv[3 mv]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + A.o_bar);
    const uint32_t bar0 = smem_u32(bars);
    // barriers: 0,1 set full | 2,3 set empty | 4 accumulators final
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
    constexpr bool gate = GATE;

    for (int t = tid; t < A.mz; t += T3_THREADS) norm[t] = A.nz ? A.nz[t] : 1.0f;
    for (int t = tid; t < 3 * A.mv; t += T3_THREADS) norm[A.mz + t] = A.nv ? A.nv[t] : 1.0f;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(BAR(i), W3);
            mbar_init(BAR(2 + i), 1);
        }
        mbar_init(BAR(4), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // zero both operand sets once: padding slots are never written again and must stay finite (0 * x)
        float4* z = reinterpret_cast<float4*>(smraw + A.o_set);
        for (int t = tid; t < (2 * SETB) >> 4; t += T3_THREADS) z[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    fence_proxy_async();
    if (warp == W3) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long long R = A.rows;
    const long long ntiles = (R + TW - 1) / TW;
    const int nt = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const uint32_t accP = 0, accD = 64, accV = 112;

    if (warp == W3) {
        // ================= MMA issuer
        const uint32_t sb = smem_u32(smraw) + A.o_set;
        const uint32_t idP = make_idesc_ex(A.MP, 64, 1, 1), idD = make_idesc_ex(64, A.N2, 1, 1), idV = make_idesc_ex(64, 16, 1, 1);
        for (int it = 0; it < nt; ++it) {
            const int b = it & 1;
            mbar_wait(BAR(b), (it >> 1) & 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t s0 = sb + (uint32_t)b * SETB;
                const uint64_t dS = mk_desc_mn(s0 + cS * CH), dD = mk_desc_mn(s0 + cD * CH), dAV = mk_desc_mn(s0 + cAV * CH);
                const uint64_t dT1 = mk_desc_mn(s0 + cT1 * CH), dT2 = mk_desc_mn(s0 + cT2 * CH), dT3 = mk_desc_mn(s0 + cT3 * CH);
                const uint64_t lo = (uint64_t)(HALFB >> 4), chs = (uint64_t)(CH >> 4);
#pragma unroll
                for (int ks = 0; ks < TW / 8; ++ks) {
                    const uint32_t acc0 = (it == 0 && ks == 0) ? 0u : 1u;
                    const uint64_t ko = (uint64_t)(ks * 64);   // 8 rows x 128 B, in 16-byte units
                    tc_mma_tf32(tmem_base + accP, dS + ko, dT1 + ko, idP, acc0);
                    tc_mma_tf32(tmem_base + accP, dS + ko, dT1 + ko + lo, idP, 1u);
                    tc_mma_tf32(tmem_base + accP, dS + ko + lo, dT1 + ko, idP, 1u);
                    tc_mma_tf32(tmem_base + accD, dD + ko, dT2 + ko, idD, acc0);
                    tc_mma_tf32(tmem_base + accD, dD + ko, dT2 + ko + lo, idD, 1u);
                    tc_mma_tf32(tmem_base + accD, dD + ko + lo, dT2 + ko, idD, 1u);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        // T3x: chunk cT3 columns 0..15, T3y: same chunk columns 16..31 (+64 B), T3z: next chunk columns 0..15
                        const uint64_t oa = ko + (uint64_t)c * chs, ob = ko + (c == 1 ? 4ull : (c == 2 ? chs : 0ull));
                        tc_mma_tf32(tmem_base + accV, dAV + oa, dT3 + ob, idV, c == 0 ? acc0 : 1u);
                        tc_mma_tf32(tmem_base + accV, dAV + oa, dT3 + ob + lo, idV, 1u);
                        tc_mma_tf32(tmem_base + accV, dAV + oa + lo, dT3 + ob, idV, 1u);
                    }
                }
                tc_commit(BAR(2 + b));
            }
            __syncwarp();
        }
        if (lane == 0) tc_commit(BAR(4));
        __syncwarp();
    } else {
        // ================= workers.  lane = (piece pc = lane & 7 : 16 bytes / one item, row-in-quad r4 = lane >> 3)
        const int pc = lane & 7, r4 = lane >> 3;
        const WarpT T = A.wt[warp];
        const float cs = A.epi.cs, cg = A.epi.cg;
        const float* nzs = norm;
        const float* nvs = norm + A.mz;
        const int nVI = A.mv >> 1;
        // ---- prefetched registers of the next tile
        float4 xsR = make_float4(0.f, 0.f, 0.f, 0.f);                 // XS: 4 scalars
        float2 xvR[3], exR = make_float2(0.f, 0.f);                   // XV: 2 vector channels (or the leftover scalar pair)
        float4 xvY = xsR;
        float2 hR[2], hvRg, hvRv[3], hvGv[3];                         // HS raw; HX raw gates / raw vectors / cotangent vectors
        float4 hG = xsR, hY = xsR;
        long long xs_i = 0, xv_i = 0, xs_in = 0, xv_in = 0, hg_i = 0, hg_in = 0;
        const SegW xsS = A.seg[T.xs_seg >= 0 ? T.xs_seg : 0], xvS = A.seg[T.xv_seg >= 0 ? T.xv_seg : 0];
        const int xs_row = T.xs_q * 4 + r4, xv_row = T.xv_q * 4 + r4;
        const int h_q = T.hs_q >= 0 ? T.hs_q : T.hx_q;
        const int h_row = h_q * 4 + r4;
        auto clampr = [&](long long gr) { return gr > R - 1 ? R - 1 : gr; };
        auto load_idx = [&](int it) {
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * TW;
            if (T.xs_seg >= 0) {
                const long long gr = clampr(row0 + xs_row);
                const int32_t* ip = xsS.idx;
                xs_in = ip ? (long long)ldgi_v(ip + gr) : gr;
            }
            if (T.xv_seg >= 0) {
                const long long gr = clampr(row0 + xv_row);
                const int32_t* ip = xvS.idx;
                xv_in = ip ? (long long)ldgi_v(ip + gr) : gr;
            }
            if (h_q >= 0) {
                const long long gr = clampr(row0 + h_row);
                hg_in = A.gout_idx ? (long long)ldgi_v(A.gout_idx + gr) : gr;
            }
        };
        auto load_rows = [&](int it) {
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * TW;
            if (T.xs_seg >= 0) {
                const SegW& S = xsS;
                if (4 * pc + 3 < S.nss) xsR = ldg4_v(S.base + xs_i * S.ld + 4 * pc);
            }
            if (T.xv_seg >= 0) {
                const SegW& S = xvS;
                const float* rp = S.base + xv_i * S.ld;
                if (pc < (S.nvs >> 1)) {
#pragma unroll
                    for (int u = 0; u < 3; ++u) xvR[u] = ldg2_v(rp + S.vcol + 6 * pc + 2 * u);
                    xvY = ldg4_v(A.in2 + clampr(row0 + xv_row) * 4);
                } else if (pc == (S.nvs >> 1) && (S.nss & 3)) {
                    xvR[0] = ldg2_v(rp + (S.nss & ~3));
                }
            }
            if (T.ex_seg >= 0) {
                const SegW& S = A.seg[T.ex_seg];
                const long long gr = clampr(row0 + lane);
                const float* rp = S.base + (S.idx ? (long long)__ldg(S.idx + gr) : gr) * S.ld;
                exR.x = __ldg(rp);
                exR.y = S.w > 1 ? __ldg(rp + 1) : 0.0f;
            }
            if (h_q >= 0) {
                const long long gr = clampr(row0 + h_row);
                hY = ldg4_v(A.in2 + gr * 4);
                if (T.hs_q >= 0) {
                    const int c = 4 * pc;
                    if (c + 3 < A.nsz) {
                        if (gate) {
                            const float* rp = A.raw + gr * A.d_out + A.oz0 + c;
                            hR[0] = ldg2_v(rp); hR[1] = ldg2_v(rp + 2);
                            hG = ldg4_v(A.gout + hg_i * A.gwidth + c);
                        } else {
                            const float* gp = A.gout + hg_i * A.gwidth + A.oz0 + c;
                            const float2 a = ldg2_v(gp), b2 = ldg2_v(gp + 2);
                            hG = make_float4(a.x, a.y, b2.x, b2.y);
                        }
                    }
                } else {
                    if (pc < nVI) {
                        if (gate) {
                            hvRg = ldg2_v(A.raw + gr * A.d_out + A.oz0 + A.nsz + 2 * pc);
                            const float* rv = A.raw + gr * A.d_out + A.ov0 + 6 * pc;
                            const float* gv = A.gout + hg_i * A.gwidth + A.nsz + 6 * pc;
#pragma unroll
                            for (int u = 0; u < 3; ++u) { hvRv[u] = ldg2_v(rv + 2 * u); hvGv[u] = ldg2_v(gv + 2 * u); }
                        } else {
                            const float* gv = A.gout + hg_i * A.gwidth + A.ov0 + 6 * pc;
#pragma unroll
                            for (int u = 0; u < 3; ++u) hvGv[u] = ldg2_v(gv + 2 * u);
                        }
                    } else if (pc == nVI && (A.nsz & 3)) {   // leftover scalar pair
                        const int c = A.nsz & ~3;
                        if (gate) {
                            hvRg = ldg2_v(A.raw + gr * A.d_out + A.oz0 + c);
                            hvGv[0] = ldg2_v(A.gout + hg_i * A.gwidth + c);
                        } else {
                            hvGv[0] = ldg2_v(A.gout + hg_i * A.gwidth + A.oz0 + c);
                        }
                    }
                }
            }
        };
        auto st4 = [&](unsigned char* set, int off, float a, float b2, float c, float d) {
            float4 h, l;
            split_tf32(a, h.x, l.x); split_tf32(b2, h.y, l.y); split_tf32(c, h.z, l.z); split_tf32(d, h.w, l.w);
            *reinterpret_cast<float4*>(set + off) = h;
            *reinterpret_cast<float4*>(set + HALFB + off) = l;
        };
        auto st2 = [&](unsigned char* set, int off, float a, float b2) {
            float2 h, l;
            split_tf32(a, h.x, l.x); split_tf32(b2, h.y, l.y);
            *reinterpret_cast<float2*>(set + off) = h;
            *reinterpret_cast<float2*>(set + HALFB + off) = l;
        };
        auto st1 = [&](unsigned char* set, int off, float a) {
            float h, l;
            split_tf32(a, h, l);
            *reinterpret_cast<float*>(set + off) = h;
            *reinterpret_cast<float*>(set + HALFB + off) = l;
        };
        auto swish_vjp = [&](float x, float g) { const float s = sigm(x); return g * cs * s * (1.0f + x * (1.0f - s)); };
        auto build = [&](int it, int b) {
            unsigned char* set = smraw + A.o_set + b * SETB;
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * TW;
            if (T.xs_seg >= 0) {
                const SegW& S = xsS;
                if (4 * pc + 3 < S.nss) st4(set, tw_off(cS, xs_row, S.koff + 4 * pc), xsR.x, xsR.y, xsR.z, xsR.w);
            }
            if (T.xv_seg >= 0) {
                const SegW& S = xvS;
                if (pc < (S.nvs >> 1)) {
                    const float v[6] = {xvR[0].x, xvR[0].y, xvR[1].x, xvR[1].y, xvR[2].x, xvR[2].y};
                    const float y1 = C3f * xvY.y, y2 = C3f * xvY.z, y3 = C3f * xvY.w, y0 = C3f * xvY.x;
                    const int kd = S.kdoff + 2 * pc;
                    st2(set, tw_off(cD, xv_row, kd), y1 * v[0] + y2 * v[1] + y3 * v[2], y1 * v[3] + y2 * v[4] + y3 * v[5]);
#pragma unroll
                    for (int c = 0; c < 3; ++c) st2(set, tw_off(cAV + c, xv_row, kd), y0 * v[c], y0 * v[3 + c]);
                } else if (pc == (S.nvs >> 1) && (S.nss & 3)) {
                    st2(set, tw_off(cS, xv_row, S.koff + (S.nss & ~3)), xvR[0].x, xvR[0].y);
                }
            }
            if (T.ex_seg >= 0) {
                st1(set, tw_off(cS, lane, A.exslot[0]), exR.x);
                if (A.seg[T.ex_seg].w > 1) st1(set, tw_off(cS, lane, A.exslot[1]), exR.y);
            }
            if (h_q >= 0) {
                const float valid = (row0 + h_row < R) ? 1.0f : 0.0f;   // rows past the end contribute nothing
                const float y0 = hY.x;
                if (T.hs_q >= 0) {
                    const int m0 = 4 * pc;
                    if (m0 + 3 < A.nsz) {
                        const float gg[4] = {hG.x, hG.y, hG.z, hG.w};
                        const float rr[4] = {hR[0].x, hR[0].y, hR[1].x, hR[1].y};
                        float h[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) h[u] = valid * nzs[m0 + u] * (gate ? swish_vjp(rr[u], gg[u]) : gg[u]);
                        st4(set, tw_off(cT2, h_row, m0), h[0], h[1], h[2], h[3]);
                        st4(set, tw_off(cT1, h_row, m0), y0 * h[0], y0 * h[1], y0 * h[2], y0 * h[3]);
                    }
                } else {
                    if (pc < nVI) {
                        const int v0 = 2 * pc;
                        const float gv[6] = {hvGv[0].x, hvGv[0].y, hvGv[1].x, hvGv[1].y, hvGv[2].x, hvGv[2].y};
                        float hv[6], hz[2] = {0.f, 0.f}, hg[2];
                        if (gate) {
                            const float rv[6] = {hvRv[0].x, hvRv[0].y, hvRv[1].x, hvRv[1].y, hvRv[2].x, hvRv[2].y};
                            const float rg[2] = {hvRg.x, hvRg.y};
#pragma unroll
                            for (int u = 0; u < 2; ++u) {
                                const float s = sigm(rg[u]);
                                const float dot = gv[3 * u] * rv[3 * u] + gv[3 * u + 1] * rv[3 * u + 1] + gv[3 * u + 2] * rv[3 * u + 2];
                                hz[u] = valid * cg * s * (1.0f - s) * dot * nzs[A.nsz + v0 + u];
                                const float sg = valid * cg * s;
#pragma unroll
                                for (int c = 0; c < 3; ++c) hv[3 * u + c] = sg * gv[3 * u + c] * nvs[3 * (v0 + u) + c];
                            }
                        } else {
#pragma unroll
                            for (int u = 0; u < 2; ++u)
#pragma unroll
                                for (int c = 0; c < 3; ++c) hv[3 * u + c] = valid * gv[3 * u + c] * nvs[3 * (v0 + u) + c];
                        }
#pragma unroll
                        for (int u = 0; u < 2; ++u) hg[u] = C3f * (hY.y * hv[3 * u] + hY.z * hv[3 * u + 1] + hY.w * hv[3 * u + 2]);
                        if (gate) {
                            st2(set, tw_off(cT2, h_row, A.nsz + v0), hz[0], hz[1]);
                            st2(set, tw_off(cT1, h_row, A.nsz + v0), y0 * hz[0], y0 * hz[1]);
                        }
                        st2(set, tw_off(cT1, h_row, 48 + v0), hg[0], hg[1]);
                        st2(set, tw_off(cT3, h_row, v0), hv[0], hv[3]);            // T3x: chunk cT3, slots 0..15
                        st2(set, tw_off(cT3, h_row, 16 + v0), hv[1], hv[4]);       // T3y: chunk cT3, slots 16..31
                        st2(set, tw_off(cT3 + 1, h_row, v0), hv[2], hv[5]);        // T3z: chunk cT3 + 1
                    } else if (pc == nVI && (A.nsz & 3)) {
                        const int m0 = A.nsz & ~3;
                        float h[2];
                        const float gg[2] = {hvGv[0].x, hvGv[0].y}, rr[2] = {hvRg.x, hvRg.y};
#pragma unroll
                        for (int u = 0; u < 2; ++u) h[u] = valid * nzs[m0 + u] * (gate ? swish_vjp(rr[u], gg[u]) : gg[u]);
                        st2(set, tw_off(cT2, h_row, m0), h[0], h[1]);
                        st2(set, tw_off(cT1, h_row, m0), y0 * h[0], y0 * h[1]);
                    }
                }
            }
        };

        if (nt > 0) {
            load_idx(0);
            xs_i = xs_in; xv_i = xv_in; hg_i = hg_in;
            load_rows(0);
            if (nt > 1) load_idx(1);
        }
        for (int it = 0; it < nt; ++it) {
            const int b = it & 1;
            mbar_wait(BAR(2 + b), ((it >> 1) & 1) ^ 1);   // the MMAs that read this set two tiles ago are done
            build(it, b);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(b));
            if (it + 1 < nt) {
                xs_i = xs_in; xv_i = xv_in; hg_i = hg_in;
                load_rows(it + 1);
                if (it + 2 < nt) load_idx(it + 2);
            }
        }
        // ---------------- final epilogue (warps 0-3): TMEM accumulators -> per-CTA partials
        if (warp < 4) {
            mbar_wait(BAR(4), 0);
            tc_fence_after();
            float* part = A.partials + (long long)blockIdx.x * A.wtot;
            const uint32_t tq = tmem_base + ((uint32_t)(32 * warp) << 16);
            // S rows: M = 128 -> TMEM lane = slot; M = 64 -> slot 16 q + i lives in lane 32 q + i
            const int slot = A.MP == 128 ? 32 * warp + lane : 16 * warp + (lane & 15);
            const bool s_ok = (A.MP == 128 || lane < 16) && slot < A.K1;
            const int ch = s_ok ? A.sl2ch[slot] : -1;
            const int kd = 16 * warp + (lane & 15);
            const bool kd_ok = lane < 16 && kd < A.nd;
            for (int m0 = 0; m0 < 64; m0 += 8) {
                float a[8];
                tc_ld8(tq + accP + m0, a);
                tc_wait_ld();
                if (ch >= 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int m = m0 + j;
                        if (m < A.mz) part[A.gw_z_off + (long long)ch * A.mz + m] = a[j];
                        else if (m >= 48 && m - 48 < A.mv) part[A.gw_v_off + (long long)ch * A.mv + (m - 48)] = a[j];
                    }
                }
            }
            for (int m0 = 0; m0 < A.N2; m0 += 8) {
                float a[8];
                tc_ld8(tq + accD + m0, a);
                tc_wait_ld();
                if (kd_ok) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (m0 + j < A.mz) part[A.gw_z_off + (long long)(A.ns + kd) * A.mz + m0 + j] = a[j];
                }
            }
            for (int m0 = 0; m0 < 16; m0 += 8) {
                float a[8];
                tc_ld8(tq + accV + m0, a);
                tc_wait_ld();
                if (kd_ok) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (m0 + j < A.mv) part[A.gw_v_off + (long long)(A.ns + kd) * A.mv + m0 + j] = a[j];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W3) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base) : "memory");
    }
}

}  // namespace se3

using namespace se3;

// weight gradients, second generation; launched=false when the configuration is not eligible (the caller falls back
// to the generic fp32 kernel)
int se3_l1tp_tc2_try_backward_w(const int n[4], const int m[4], const int t_in[4], const int t_out[4], int ntab,
                                const int* h_tab, const se3_l1tp_bwd_args* a, const RowSrc& src, const EpiL& epi,
                                float* partials, int wtot, int gw_z_off, int gw_v_off, int max_grid, cudaStream_t st,
                                int* grid_out, bool* launched) {
    *launched = false;
    static int disabled = -1;
    if (disabled < 0) {
        const char* e = getenv("SE3_DISABLE_TC2");
        const char* e1 = getenv("SE3_DISABLE_TC");
        disabled = ((e && (e[0] == '1' || e[0] == '2')) || (e1 && (e1[0] == '1' || e1[0] == '2'))) ? 1 : 0;
    }
    if (disabled) return SE3_OK;
    if (n[1] || n[2] || m[1] || m[2]) return SE3_OK;
    const int ns = n[0], nd = n[3], mz = m[0], mv = m[3];
    if (ns < 1 || nd < 1 || mz < 1 || mv < 1) return SE3_OK;
    if (a->rows >= (1ll << 31) - TW) return SE3_OK;
    static thread_local Tc3Args A;
    memset(&A, 0, sizeof(A));
    const bool gate = epi.mode == SE3_EPI_GATE;
    A.ns = ns; A.nd = nd; A.mz = mz; A.mv = mv; A.d_out = mz + 3 * mv;
    A.oz0 = h_tab[t_out[0]]; A.ov0 = h_tab[t_out[3]];
    for (int k = 0; k < mz; ++k) if (h_tab[t_out[0] + k] != A.oz0 + k) return SE3_OK;
    for (int k = 0; k < mv; ++k) if (h_tab[t_out[3] + k] != A.ov0 + 3 * k) return SE3_OK;
    A.nsz = gate ? epi.ns_g : mz;
    if (gate && (epi.ns_g < 1 || epi.ns_g + mv != mz)) return SE3_OK;
    A.gwidth = epi.d_post;
    if ((mv & 1) || (A.nsz & 1) || (A.oz0 & 1) || (A.ov0 & 1) || (A.d_out & 1) || (A.gwidth & 1)) return SE3_OK;
    if (((uintptr_t)a->in2 & 15) || ((uintptr_t)a->gout & 15) || (gate && ((uintptr_t)a->raw & 7))) return SE3_OK;
    if (gate && (A.gwidth & 3)) return SE3_OK;
    if (mz > 48 || mv > 16 || (mv >> 1) + ((A.nsz & 3) ? 1 : 0) > 8 || (A.nsz >> 2) > 8) return SE3_OK;
    A.N2 = (mz + 15) & ~15; A.N3 = 16;
    // ---- segments: wide = [nss scalars][3 nvs vector components], 16-byte rows; at most one narrow (<= 2 scalars) segment
    std::vector<int> kind(src.cum[src.nseg], 0), chan(src.cum[src.nseg], -1);
    for (int k = 0; k < ns; ++k) { const int c = h_tab[t_in[0] + k]; kind[c] = 1; chan[c] = k; }
    for (int k = 0; k < nd; ++k) { const int c = h_tab[t_in[3] + k]; for (int q = 0; q < 3; ++q) { kind[c + q] = 2 + q; chan[c + q] = k; } }
    for (int k = 0; k < 96; ++k) A.sl2ch[k] = -1;
    int K1 = 0, kd = 0, exseg = -1;
    std::vector<int> freeslots;
    A.nseg = src.nseg;
    for (int s = 0; s < src.nseg; ++s) {
        SegW& S = A.seg[s];
        const int w = src.cum[s + 1] - src.cum[s], c0 = src.cum[s];
        S.base = src.base[s]; S.idx = src.idx[s]; S.ld = src.ld[s]; S.w = w;
        S.wide = ((w & 3) == 0 && (src.ld[s] & 3) == 0 && ((uintptr_t)src.base[s] & 15) == 0) ? 1 : 0;
        if (!S.wide) {
            if (exseg >= 0 || w > 2) return SE3_OK;
            for (int c = 0; c < w; ++c) if (kind[c0 + c] != 1) return SE3_OK;
            exseg = s;
            continue;
        }
        int nss = 0;
        while (nss < w && kind[c0 + nss] == 1) ++nss;
        const int nv3 = w - nss;
        if (nv3 % 3) return SE3_OK;
        for (int c = nss; c < w; ++c) if (kind[c0 + c] != 2 + (c - nss) % 3) return SE3_OK;
        S.nss = nss; S.nvs = nv3 / 3; S.vcol = nss;
        if ((nss & 1) || (S.nvs & 1) || (nss >> 2) > 8 || (S.nvs >> 1) + ((nss & 3) ? 1 : 0) > 8) return SE3_OK;
        // scalar channels of a segment are consecutive in the species order: slot = koff + local scalar index
        for (int c = 0; c < nss; ++c) if (chan[c0 + c] != chan[c0] + c) return SE3_OK;
        for (int c = 0; c < S.nvs; ++c) if (chan[c0 + nss + 3 * c] != chan[c0 + nss] + c) return SE3_OK;
        S.koff = K1;
        for (int c = 0; c < nss; ++c) A.sl2ch[K1 + c] = (short)chan[c0 + c];
        const int span = (nss + 3) & ~3;
        for (int c = nss; c < span; ++c) freeslots.push_back(K1 + c);
        K1 += span;
        if (S.nvs) {
            if (chan[c0 + nss] != kd) return SE3_OK;   // vector channels in order across segments
            S.kdoff = kd;
            kd += S.nvs;
        }
        if (K1 > 96) return SE3_OK;
    }
    if (kd != nd || kd > 32) return SE3_OK;
    if (exseg >= 0) {
        const int w = A.seg[exseg].w, c0 = src.cum[exseg];
        for (int c = 0; c < w; ++c) {
            int sl;
            if ((size_t)c < freeslots.size()) sl = freeslots[c];
            else { sl = K1++; if (K1 > 96) return SE3_OK; }
            A.exslot[c] = (short)sl;
            A.sl2ch[sl] = (short)chan[c0 + c];
        }
    }
    {   // every scalar channel must have a slot
        int cnt = 0;
        for (int k = 0; k < 96; ++k) cnt += A.sl2ch[k] >= 0;
        if (cnt != ns) return SE3_OK;
    }
    A.K1 = K1; A.MP = K1 > 64 ? 128 : 64;
    // ---- tasks: one XS / XV task per (wide segment, row quad), 8 HS + 8 HX tasks; each warp takes at most one per kind
    for (int w = 0; w < W3; ++w) { WarpT& T = A.wt[w]; T.xs_seg = T.xs_q = T.xv_seg = T.xv_q = T.hs_q = T.hx_q = T.ex_seg = -1; }
    struct Tk { int kind, seg, q, cost; };
    std::vector<Tk> tasks;
    for (int s = 0; s < src.nseg; ++s) {
        if (!A.seg[s].wide) { tasks.push_back({4, s, 0, 10}); continue; }
        for (int q = 0; q < TW / 4; ++q) {
            if (A.seg[s].nss >= 4) tasks.push_back({0, s, q, 15});
            if (A.seg[s].nvs || (A.seg[s].nss & 3)) tasks.push_back({1, s, q, 70});
        }
    }
    for (int q = 0; q < TW / 4; ++q) {
        if (A.nsz >= 4) tasks.push_back({2, -1, q, 80});
        tasks.push_back({3, -1, q, 110});
    }
    std::stable_sort(tasks.begin(), tasks.end(), [](const Tk& x, const Tk& y) { return x.cost > y.cost; });
    int load[W3] = {0};
    for (const Tk& t : tasks) {
        int best = -1;
        for (int w = 0; w < W3; ++w) {
            const WarpT& T = A.wt[w];
            const bool free_ = t.kind == 0 ? T.xs_seg < 0 : t.kind == 1 ? T.xv_seg < 0 : t.kind == 4 ? T.ex_seg < 0
                               : (T.hs_q < 0 && T.hx_q < 0);   // one H task per warp: they share the prefetch registers
            if (free_ && (best < 0 || load[w] < load[best])) best = w;
        }
        if (best < 0) return SE3_OK;
        WarpT& T = A.wt[best];
        if (t.kind == 0) { T.xs_seg = (short)t.seg; T.xs_q = (short)t.q; }
        else if (t.kind == 1) { T.xv_seg = (short)t.seg; T.xv_q = (short)t.q; }
        else if (t.kind == 2) T.hs_q = (short)t.q;
        else if (t.kind == 3) T.hx_q = (short)t.q;
        else T.ex_seg = (short)t.seg;
        load[best] += t.cost;
    }
    A.rows = a->rows; A.in2 = a->in2; A.nz = a->norm[0]; A.nv = a->norm[3];
    A.epi = epi; A.raw = a->raw; A.gout = a->gout; A.gout_idx = a->gout_idx;
    A.partials = partials; A.wtot = wtot; A.gw_z_off = gw_z_off; A.gw_v_off = gw_v_off;
    auto al = [](int x, int q) { return (x + q - 1) / q * q; };
    int o = 0;
    A.o_set = o; o += 2 * SETB;
    A.o_norm = o; o += al((mz + 3 * mv) * 4, 16);
    A.o_bar = o; o += 8 * 8 + 16;
    int dev = 0, maxsm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (o > maxsm) return SE3_OK;
    static bool attr_set = false;
    if (!attr_set) {
        SE3_CUDA_TRY(cudaFuncSetAttribute(l1tp_tc2_bwdw_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
        SE3_CUDA_TRY(cudaFuncSetAttribute(l1tp_tc2_bwdw_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
        attr_set = true;
    }
    const long long ntiles = (a->rows + TW - 1) / TW;
    const int grid = (int)std::min<long long>(ntiles, std::min(num_sms(), max_grid));
    if (gate) l1tp_tc2_bwdw_kernel<true><<<grid, T3_THREADS, o, st>>>(A);
    else l1tp_tc2_bwdw_kernel<false><<<grid, T3_THREADS, o, st>>>(A);
    SE3_LAUNCHED();
    g_tc_launches.fetch_add(1, std::memory_order_relaxed);
    *grid_out = grid;
    *launched = true;
    return SE3_OK;
}
