// Weight gradient of the l <= 2 tensor product on the tensor cores (sm_100a: tcgen05.mma kind::tf32 as 3xTF32,
// accumulators resident in TMEM for all tiles of a CTA) — BASELINE configs[2]; the l <= 1 counterparts are
// l1tp_tc2_bwdw.cu / msg_fused_bwdw.cu, whose operand layout and skeleton this kernel shares.
//
// Formulation ("contract first, couple after", transposed).  With g' = a_io g the cotangent and M_p(Y)[i][c] the coupling
// of path p = (i1, i2, io) folded with the row's second input (o3tp_cg_gen.inl),
//     GT[e][col_p + w d1 + i] = sum_c M_p(Y_e)[i][c] g'[e][io][w][c]            (SIMT, lane = row, 4 channels per item)
//     gW_p[u][w]              = sum_e sum_i x[e][off1_p + u d1 + i] GT[e][col_p + w d1 + i]
// so ONE accumulation P = X^T GT over the rows (M = d_in1 <= 64 columns of x, N = all GT columns <= 512, K = rows) holds
// every weight gradient on the "diagonals" of its blocks; the cross terms are computed and never read — the tensor pipe
// is idle anyway, what counts is that the 7 204 multiply-adds per row of the SIMT kernel pair become 6 tcgen05.mma per
// 8 rows.  The rows are the MMA K dimension: both row-major tiles are MN-major operands in the SWIZZLE_128B_BASE32B
// layout of l1tp_tc2_bwdw.cu (128-byte column chunks of 32 slots, 4-row atoms, 32-byte units XOR-ed with row & 3).
//
// Feeding: the tile's rows of in1, of the cotangent and of in2 are contiguous in HBM: three cp.async.bulk copies per
// 32-row tile into a double-buffered staging area, issued two tiles ahead.  The GT operand (122 KB hi / lo for the
// message product) exists once, so it is built and consumed in TWO phases (two groups of paths = two column ranges of
// <= 256 columns, one MMA per K-step each): while the MMAs of phase A run the workers build phase B, and the MMAs of
// phase B run under the next tile's phase A; the small x operand is double buffered.
// Only whole 32-row tiles are taken; the caller runs the SIMT kernel on the remaining rows.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "o3tp_tc.h"
#include "tc_common.cuh"

namespace {

using namespace se3;

#define O3_DEV __device__ __forceinline__
#include "o3tp_cg_gen.inl"

constexpr int GW = 16;                       // worker warps
constexpr int G_THREADS = (GW + 1) * 32, GWT = GW * 32;
constexpr int TW = 32;                       // rows per tile (4 MMA K-steps)
constexpr int CHB = TW * 128;                // bytes of one 32-slot chunk
constexpr int MAXPATH = 32, MAXIO = 8, MAXITEM = 256, CH = 4;

struct PathD { int off1, l1, mul1, l2, yoff, io, colbase, woff; };
struct IoD { int off, mul, l; float a; };
struct Tab {
    int npath, nio, nitem, D1, D2, Dout, nX, nG;   // nX / nG: 32-slot chunks of the x / GT operand
    PathD path[MAXPATH];
    IoD io[MAXIO];
    int item[MAXITEM];                             // path | first channel << 8, heaviest first inside a phase
    int nphase, ph_chunk[3], ph_item[3];           // phase p: GT chunks [ph_chunk[p], ph_chunk[p+1]), items [ph_item[p], ph_item[p+1])
};

struct GwArgs {
    Tab T;
    long long rows;                                // multiple of TW
    const float* x;
    const float* y;
    const float* g;
    float* partials;                               // [grid][64][32 nG]
    int o_x, o_g, o_y, xb, gb, yb, o_tab, o_bar;
    int xhalf, xset, o_gt, ghalf;                  // bytes: hi part of an x set, one x set (hi | lo), GT hi offset, GT hi part
    int nstage;                                    // staging buffers (2; 1 when the cotangent rows are too wide for two)
};

__device__ __forceinline__ int tw_off(int chunk, int row, int slot) {   // byte offset inside the hi part of the set
    return chunk * CHB + row * 128 + ((((slot >> 3)) ^ (row & 3)) << 5) + ((slot & 7) << 2);
}
__device__ __forceinline__ uint64_t mk_desc_mn(uint32_t saddr) {        // MN-major, SWIZZLE_128B_BASE32B
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((CHB >> 4) & 0x3FFF) << 16) | ((uint64_t)((512 >> 4) & 0x3FFF) << 32) |
           (1ull << 46) | (1ull << 61);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st4(unsigned char* set, int off, int half, float a, float b, float c, float d) {
    float4 h, l;
    split_tf32(a, h.x, l.x); split_tf32(b, h.y, l.y); split_tf32(c, h.z, l.z); split_tf32(d, h.w, l.w);
    *reinterpret_cast<float4*>(set + off) = h;
    *reinterpret_cast<float4*>(set + off + half) = l;
}

// coupling^T of four cotangent channels of one path for the row of this lane -> 4 d1 consecutive GT columns
template <int L1, int L2, int LO>
O3_DEV void gt_path(const float* __restrict__ yr, const float (&g)[CH][2 * LO + 1], unsigned char* set, int half, int chunk0,
                    int row, int col) {
    constexpr int D1 = 2 * L1 + 1, DO = 2 * LO + 1;
    float M[D1][DO];
    o3_M<L1, L2, LO>(yr, M);
    float v[CH * D1];
#pragma unroll
    for (int k = 0; k < CH; ++k)
#pragma unroll
        for (int i = 0; i < D1; ++i) {
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < DO; ++c)
                if ((o3_nz<L1, L2, LO>::mask >> (i * DO + c)) & 1u) s = fmaf(M[i][c], g[k][c], s);
            v[k * D1 + i] = s;
        }
#pragma unroll
    for (int j = 0; j < D1; ++j) {
        const int cc = col + 4 * j;
        st4(set, tw_off(chunk0 + (cc >> 5), row, cc & 31), half, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
}

template <int LO>
O3_DEV void gt_item(const PathD& P, const IoD& I, int w0, const float* grow, const float* yr, unsigned char* set, int half,
                    int chunk0, int row) {
    constexpr int DO = 2 * LO + 1;
    const int nw = min(CH, I.mul - w0);
    float g[CH][DO];
#pragma unroll
    for (int k = 0; k < CH; ++k)
#pragma unroll
        for (int c = 0; c < DO; ++c) g[k][c] = k < nw ? I.a * grow[I.off + (w0 + k) * DO + c] : 0.f;
    const int col = P.colbase + w0 * (2 * P.l1 + 1);
    switch (P.l1 * 9 + P.l2 * 3 + LO) {
#define O3G_CASE(a, b, c)                                                         \
    case a * 9 + b * 3 + c:                                                       \
        if constexpr (c == LO) gt_path<a, b, c>(yr + P.yoff, g, set, half, chunk0, row, col); \
        break;
        O3_TRIPLES(O3G_CASE)
#undef O3G_CASE
    }
}

__global__ void __launch_bounds__(G_THREADS, 1) o3tp_tc_gw_kernel(const __grid_constant__ GwArgs A) {
    extern __shared__ __align__(1024) unsigned char smraw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    Tab& T = *reinterpret_cast<Tab*>(smraw + A.o_tab);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + A.o_bar);
    const uint32_t bar0 = smem_u32(bars);
    // barriers: 0,1 phase operand full | 2,3 MMAs of the phase done (its GT columns free) | 4,5 staged rows landed | 6 final
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    int* ctr = reinterpret_cast<int*>(bars + 9);   // item counters [phase][tile parity]

    for (int i = tid; i < (int)(sizeof(Tab) / 4); i += G_THREADS) reinterpret_cast<int*>(&T)[i] = reinterpret_cast<const int*>(&A.T)[i];
    if (tid == 0) {
        mbar_init(BAR(0), GW);
        mbar_init(BAR(1), GW);
        for (int i = 2; i < 7; ++i) mbar_init(BAR(i), 1);
        ctr[0] = ctr[1] = ctr[2] = ctr[3] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // zero the operands once: unused slots are never written again and must stay finite (0 * x)
        float4* z = reinterpret_cast<float4*>(smraw);
        for (int t = tid; t < (2 * A.xset + 2 * A.ghalf) >> 4; t += G_THREADS) z[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    fence_proxy_async();
    if (warp == GW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long long ntiles = A.rows / TW;
    const int nt = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const int nG = A.T.nG, nph = A.T.nphase;

    if (warp == GW) {
        // ================= MMA issuer: P[64 x 32 nG] += X^T GT, one phase (column range) after the other
        const uint32_t sb = smem_u32(smraw);
        const uint64_t lox = (uint64_t)(A.xhalf >> 4), log = (uint64_t)(A.ghalf >> 4);
        for (int it = 0; it < nt; ++it) {
            const uint64_t dX = mk_desc_mn(sb + (uint32_t)(it & 1) * A.xset);
            for (int ph = 0; ph < nph; ++ph) {
                mbar_wait(BAR(ph), it & 1);
                tc_fence_after();
                if (lane == 0) {
                    const int cb = A.T.ph_chunk[ph], ce = A.T.ph_chunk[ph + 1];
                    for (int ks = 0; ks < TW / 8; ++ks) {
                        const uint32_t acc0 = (it == 0 && ks == 0) ? 0u : 1u;
                        const uint64_t ko = (uint64_t)(ks * 64);   // 8 rows x 128 B, in 16-byte units
                        for (int c0 = cb; c0 < ce; c0 += 8) {
                            const int nc = min(8, ce - c0);
                            const uint32_t id = make_idesc_ex(64, 32 * nc, 1, 1);
                            const uint64_t dG = mk_desc_mn(sb + A.o_gt + (uint32_t)c0 * CHB);
                            const uint32_t d = tmem_base + 32u * c0;
                            tc_mma_tf32(d, dX + ko, dG + ko, id, acc0);
                            tc_mma_tf32(d, dX + ko, dG + ko + log, id, 1u);
                            tc_mma_tf32(d, dX + ko + lox, dG + ko, id, 1u);
                        }
                    }
                    tc_commit(BAR(2 + ph));
                }
                __syncwarp();
            }
        }
        if (lane == 0) tc_commit(BAR(6));
        __syncwarp();
    } else {
        // ================= workers
        const uint32_t sm_u32 = smem_u32(smraw);
        const int ns = A.nstage;
        auto issue_pf = [&](int it) {      // warp 0, lane 0: the three contiguous blocks of tile `it` -> staging buffer it % ns
            const long long row0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * TW;
            const int b = ns == 2 ? (it & 1) : 0;
            const uint32_t bx = TW * T.D1 * 4, bg = TW * T.Dout * 4, by = TW * T.D2 * 4;
            mbar_arrive_tx(BAR(4 + b), bx + bg + by);
            bulk_g2s(sm_u32 + A.o_x + b * A.xb, A.x + row0 * T.D1, bx, BAR(4 + b));
            bulk_g2s(sm_u32 + A.o_g + b * A.gb, A.g + row0 * T.Dout, bg, BAR(4 + b));
            bulk_g2s(sm_u32 + A.o_y + b * A.yb, A.y + row0 * T.D2, by, BAR(4 + b));
        };
        if (nt > 0 && warp == 0 && lane == 0) {
            issue_pf(0);
            if (nt > 1 && ns == 2) issue_pf(1);
        }
        const int xpieces = T.D1 >> 2;     // 16-byte pieces per x row
        unsigned char* gt = smraw + A.o_gt;
        for (int it = 0; it < nt; ++it) {
            const int b = it & 1, sb = ns == 2 ? b : 0;
            if (tid == 0) { ctr[b ^ 1] = 0; ctr[2 + (b ^ 1)] = 0; }   // the other parity: last used a tile ago, next used a tile ahead
            mbar_wait(BAR(4 + sb), ns == 2 ? (it >> 1) & 1 : it & 1);   // staged rows of this tile landed
            const float* xs = reinterpret_cast<const float*>(smraw + A.o_x + sb * A.xb);
            const float* gs = reinterpret_cast<const float*>(smraw + A.o_g + sb * A.gb);
            const float* ys = reinterpret_cast<const float*>(smraw + A.o_y + sb * A.yb);
            for (int ph = 0; ph < nph; ++ph) {
                if (it > 0) mbar_wait(BAR(2 + ph), (it - 1) & 1);    // the MMAs that read this phase's GT columns are done
                if (ph == 0) {
                    // x rows -> M-side operand (set it & 1: its readers, the MMAs of tile it - 2, are long done): task =
                    // (piece, row); a warp covers 8 pieces x 4 rows of one atom (conflict-free)
                    unsigned char* xset = smraw + b * A.xset;
                    for (int task = tid; task < TW * 8 * ((xpieces + 7) >> 3); task += GWT) {
                        const int pc = task & 7, r4 = (task >> 3) & 3, rq = (task >> 5) & 7, piece = (task >> 8) * 8 + pc, row = rq * 4 + r4;
                        if (piece < xpieces) {
                            const float4 v = *reinterpret_cast<const float4*>(xs + row * T.D1 + 4 * piece);
                            st4(xset, tw_off(piece >> 3, row, 4 * (piece & 7)), A.xhalf, v.x, v.y, v.z, v.w);
                        }
                    }
                }
                // cotangent rows -> GT (N-side operand): lane = row, warps pull (path, 4 channels) items heaviest first
                const int i0 = T.ph_item[ph], i1 = T.ph_item[ph + 1];
                for (;;) {
                    int u = 0;
                    if (lane == 0) u = atomicAdd(&ctr[2 * ph + b], 1);
                    u = __shfl_sync(0xffffffffu, u, 0) + i0;
                    if (u >= i1) break;
                    const int code = T.item[u];
                    const PathD& P = T.path[code & 255];
                    const IoD& I = T.io[P.io];
                    const int w0 = code >> 8;
                    const float *grow = gs + lane * T.Dout, *yr = ys + lane * T.D2;
                    switch (I.l) {
                        case 0: gt_item<0>(P, I, w0, grow, yr, gt, A.ghalf, 0, lane); break;
                        case 1: gt_item<1>(P, I, w0, grow, yr, gt, A.ghalf, 0, lane); break;
                        default: gt_item<2>(P, I, w0, grow, yr, gt, A.ghalf, 0, lane); break;
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(ph));
            }
            named_bar(1, GWT);                                   // every worker is done with this staging buffer
            if (warp == 0 && lane == 0 && it + ns < nt) issue_pf(it + ns);
        }
        // ---------------- final epilogue (warps 0-3): TMEM -> this CTA's partial [64][32 nG]; M = 64: row 16 q + i lives in
        // TMEM lane 32 q + i
        if (warp < 4) {
            mbar_wait(BAR(6), 0);
            tc_fence_after();
            const int ntp = 32 * nG;
            float* part = A.partials + (long long)blockIdx.x * 64 * ntp + (long long)(16 * warp + (lane & 15)) * ntp;
            const uint32_t tq = tmem_base + ((uint32_t)(32 * warp) << 16);
            for (int c0 = 0; c0 < ntp; c0 += 8) {
                float a[8];
                tc_ld8(tq + c0, a);
                tc_wait_ld();
                if (lane < 16) {
                    reinterpret_cast<float4*>(part + c0)[0] = make_float4(a[0], a[1], a[2], a[3]);
                    reinterpret_cast<float4*>(part + c0)[1] = make_float4(a[4], a[5], a[6], a[7]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == GW) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// gw[woff_p + u mul_out + w] = sum over the CTAs' partials and over the d1 diagonal entries of the (u, w) block: one warp
// per weight, lane l sums partials l, l + 32, ... in a fixed order, then a shuffle tree (run-to-run deterministic)
__global__ void __launch_bounds__(256) o3tp_tc_gw_reduce_kernel(const __grid_constant__ Tab T, const float* __restrict__ part,
                                                                int nparts, int nw_total, float* __restrict__ gw) {
    const int lane = threadIdx.x & 31, ntp = 32 * T.nG;
    for (int t = blockIdx.x * 8 + (threadIdx.x >> 5); t < nw_total; t += gridDim.x * 8) {
        int p = 0;
        for (int q = 1; q < T.npath; ++q)
            if (t >= T.path[q].woff) p = q;
        const PathD P = T.path[p];
        const int mo = T.io[P.io].mul, d1 = 2 * P.l1 + 1;
        const int u = (t - P.woff) / mo, w = (t - P.woff) - u * mo;
        float s = 0.f;
        for (int i = 0; i < d1; ++i) {
            const float* q0 = part + (long long)(P.off1 + u * d1 + i) * ntp + P.colbase + w * d1 + i;
            for (int q = lane; q < nparts; q += 32) s += q0[(long long)q * 64 * ntp];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) gw[t] = s;
    }
}

}  // namespace

struct O3TcGw {
    GwArgs A;
    size_t smem = 0;
    int nw = 0;
    float* partials = nullptr;
    int parts = 0;
};

O3TcGw* o3tp_tc_gw_create(const o3::Plan& P) {
    if (getenv("SE3_O3TP_NO_TC")) return nullptr;
    if (P.D1 > 64 || (P.D1 & 3) || P.D2 > 32 || P.D2 < 1 || (int)P.paths.size() > MAXPATH || (int)P.out.size() > MAXIO) return nullptr;
    O3TcGw* S = new O3TcGw();
    Tab& T = S->A.T;
    memset(&T, 0, sizeof(T));
    std::vector<int> off1, off2, offo;
    int acc = 0;
    for (auto& ir : P.in1) { off1.push_back(acc); acc += ir.mul * (2 * ir.l + 1); }
    acc = 0;
    for (auto& ir : P.in2) { off2.push_back(acc); acc += 2 * ir.l + 1; }
    acc = 0;
    for (auto& ir : P.out) { offo.push_back(acc); acc += ir.mul * (2 * ir.l + 1); }
    T.npath = (int)P.paths.size(); T.nio = (int)P.out.size(); T.D1 = P.D1; T.D2 = P.D2; T.Dout = P.Dout;
    for (size_t io = 0; io < P.out.size(); ++io) T.io[io] = {offo[io], P.out[io].mul, P.out[io].l, P.a[io]};
    // two phases = two groups of paths (greedy split by padded column count, heaviest path first); each phase's columns
    // are contiguous and start on a chunk boundary
    struct Pw { int p, width; };
    std::vector<Pw> pw;
    for (int p = 0; p < T.npath; ++p) {
        const o3::PathH& h = P.paths[p];
        pw.push_back({p, ((P.out[h.io].mul + CH - 1) / CH) * CH * (2 * P.in1[h.i1].l + 1)});
    }
    std::vector<Pw> byw = pw;
    std::stable_sort(byw.begin(), byw.end(), [](const Pw& x, const Pw& y) { return x.width > y.width; });
    std::vector<int> phase_of(T.npath, 0);
    int load[2] = {0, 0};
    for (const Pw& q : byw) {
        const int k = load[1] < load[0] ? 1 : 0;
        phase_of[q.p] = k;
        load[k] += q.width;
    }
    int nphase = (load[0] > 0 && load[1] > 0 && T.npath >= 2) ? 2 : 1;
    if (nphase == 2 && ((load[0] + 31) / 32 + (load[1] + 31) / 32) * 32 > 512) nphase = 1;   // the chunk padding does not fit TMEM
    if (nphase == 1) std::fill(phase_of.begin(), phase_of.end(), 0);
    struct It { int code; double cost; };
    int col = 0, nitem = 0;
    T.nphase = nphase;
    for (int ph = 0; ph < nphase; ++ph) {
        col = (col + 31) & ~31;
        T.ph_chunk[ph] = col / 32;
        T.ph_item[ph] = nitem;
        std::vector<It> items;
        for (int p = 0; p < T.npath; ++p) {
            if (phase_of[p] != ph) continue;
            const o3::PathH& h = P.paths[p];
            const o3::Irrep a = P.in1[h.i1], b = P.in2[h.i2], o = P.out[h.io];
            const int d1 = 2 * a.l + 1, dout = 2 * o.l + 1;
            T.path[p] = {off1[h.i1], a.l, a.mul, b.l, off2[h.i2], h.io, col, h.woff};
            for (int w0 = 0; w0 < o.mul; w0 += CH) {
                const double cost = 30 + d1 * dout * (a.l && b.l && o.l ? 3.0 : 1.0) + CH * (dout + 2.0 * d1 * dout + 12.0 * d1);
                items.push_back({p | (w0 << 8), cost});
            }
            col += pw[p].width;    // 4-channel items write whole 16-byte pieces: the path's width is padded to 4 channels
        }
        std::stable_sort(items.begin(), items.end(), [](const It& x, const It& y) { return x.cost > y.cost; });
        if (nitem + (int)items.size() > MAXITEM) { delete S; return nullptr; }
        for (const It& it : items) T.item[nitem++] = it.code;
    }
    const int nG = (col + 31) / 32;
    T.ph_chunk[nphase] = nG;
    T.ph_item[nphase] = nitem;
    if (nG * 32 > 512 || T.npath > 255) { delete S; return nullptr; }
    T.nitem = nitem;
    T.nX = 2; T.nG = nG;
    GwArgs& A = S->A;
    A.xhalf = T.nX * CHB; A.xset = 2 * A.xhalf; A.o_gt = 2 * A.xset; A.ghalf = T.nG * CHB;
    auto r128 = [](int x) { return (x + 127) & ~127; };
    A.xb = r128(TW * T.D1 * 4); A.gb = r128(TW * T.Dout * 4); A.yb = r128(TW * T.D2 * 4);
    int dev = 0, maxsm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    for (A.nstage = 2; A.nstage >= 1; --A.nstage) {
        const int n = A.nstage;
        A.o_x = A.o_gt + 2 * A.ghalf; A.o_g = A.o_x + n * A.xb; A.o_y = A.o_g + n * A.gb;
        A.o_tab = A.o_y + n * A.yb; A.o_bar = A.o_tab + r128((int)sizeof(Tab));
        S->smem = (size_t)A.o_bar + 128;
        if ((int)S->smem <= maxsm) break;
    }
    S->nw = P.nW;
    if (A.nstage < 1) { delete S; return nullptr; }
    return S;
}

void o3tp_tc_gw_destroy(O3TcGw* S) {
    if (!S) return;
    if (S->partials) cudaFree(S->partials);
    delete S;
}

// gw (overwritten) = weight gradient over the first `rows` rows (a multiple of 32); returns SE3_OK or an error code
int o3tp_tc_gw_run(O3TcGw* S, long long rows, const float* x, const float* y, const float* g, float* gw, cudaStream_t st) {
    if (!S || rows <= 0 || (rows % TW) != 0) { set_error("o3tp_tc_gw: rows must be a positive multiple of %d", TW); return SE3_ERR_INVALID; }
    const int ntp = 32 * S->A.T.nG;
    // at least 120 KB of dynamic shared memory: one CTA per SM (it allocates all 512 TMEM columns)
    const size_t launch_smem = std::max<size_t>(S->smem, 120 * 1024);
    if (!S->partials) {
        S->parts = num_sms();
        SE3_CUDA_TRY(cudaMalloc(&S->partials, sizeof(float) * (size_t)S->parts * 64 * ntp));
    }
    static size_t attr = 0;
    if (launch_smem > attr) {
        SE3_CUDA_TRY(cudaFuncSetAttribute(o3tp_tc_gw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)launch_smem));
        attr = launch_smem;
    }
    GwArgs A = S->A;
    A.rows = rows; A.x = x; A.y = y; A.g = g; A.partials = S->partials;
    const int grid = (int)std::min<long long>(rows / TW, S->parts);
    o3tp_tc_gw_kernel<<<grid, G_THREADS, launch_smem, st>>>(A);
    SE3_LAUNCHED();
    g_tc_launches.fetch_add(1, std::memory_order_relaxed);
    o3tp_tc_gw_reduce_kernel<<<num_sms(), 256, 0, st>>>(A.T, S->partials, grid, S->nw, gw);
    SE3_LAUNCHED();
    return SE3_OK;
}

bool o3tp_tc_gw_aligned(const O3TcGw* S, const float* x, const float* y, const float* g) {
    return S && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(g)) & 15) == 0;
}
