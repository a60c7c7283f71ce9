// Tensor-core (tcgen05 / TMEM, 3xTF32) backward of the fused l<=1 tensor-product layer, SEGNN case
// (inputs a x0e + b x1o, outputs c x0e + d x1o).  Two kernels share the producer and the H-tile builders:
//
//  (W) l1tp_tc_bwdw_kernel — weight gradients.  The reduction runs over rows, so rows are the MMA K dimension:
//      feature tiles A[row][feature] and cotangent tiles H[row][channel] (both stored as 8-row x 16-byte core
//      matrices) are used as MN-major operands, and the four accumulators
//          gWZ_s += (S)^T (Y0 HZ)   gWV_s += S^T HG   gWZ_d += Dd^T HZ   gWV_v += sum_c AVc^T HVc
//      stay resident in TMEM across ALL tiles of the CTA (128 columns); they are read once at the end and written
//      as per-CTA partials that the deterministic reduction kernel of l1tp.cu sums.
//  (I) l1tp_tc_bwdi_kernel — input gradients.  Per 64-row tile
//          gS = [Y0 HZ | HG] . [WZ_s | WV_s]^T     gD = HZ . WZ_d^T     gTc = HVc . WV_v^T
//      then g_s = gS, g_v[c] = c3 Y1[c] gD + c3 Y0 gTc, scattered per segment (store / red.v4 / run-length sorted).
//
// H tiles (gate VJP fused): HZ = norm * g_raw(l=0), HVc = norm * g_raw(l=1)[c], HG = c3 sum_c Y1[c] HVc.
#include <algorithm>

#include "tc_common.cuh"

namespace se3 {

static constexpr int TBW_THREADS = 768;  // warps: 0,2 producers | 1 MMA | 4-15 builders (4-7 also final epilogue)
static constexpr int TMB = 32;           // rows per tile of the weight-gradient kernel (= 4 MMA K-steps)
static constexpr int BW_BUILD_W0 = 4, BW_NBUILD = 20, BW_NSLOT = 3;

struct TcBwdArgs {
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
    float* partials;
    int wtot, gw_z_off, gw_v_off;
    const int* tab;
    int ntab;
    int ns, nd, mz, mv, NSG, NDG, NSG8, NDG8, N2, N3, MS;  // MS: MMA M of the S operand (64 or 128)
    int t_s, t_d, t_oz, t_ov, d_out, gwidth;
    int sstride[SE3_MAX_SEG], soff[SE3_MAX_SEG], vec16[SE3_MAX_SEG], swidth[SE3_MAX_SEG];
    int in2off, rawoff, rawstride, goff, gstride, slot_floats;
    int graw_vec16, g_vec16, tmem_cols, use_tma;
    // shared memory byte offsets
    int o_as, o_ad, o_av, o_t1, o_t2, o_t3, o_stage, o_tab, o_norm, o_tbl, o_bar, o_b1, o_b2, o_b3, o_gt;
    int NS8, ND8, gts;
    int rg_s, rg_d, rg1, rg2, rg3;           // bytes per 8-row group of each tile
    int sz_as, sz_ad, sz_t1, sz_t2, sz_t3;   // bytes of one (hi or lo) tile
};

// ---- cotangent (H) tiles for `nrows` rows of one stage slot: warp-task = 16 rows x 2 channel groups
//   T1 = [Y0*HZ | HG]   T2 = HZ   T3[c] = HVc      element (r, ch) at (r>>3)*RG + (ch>>2)*128 + (r&7)*16 + (ch&3)*4
__device__ __forceinline__ void put4(unsigned char* tile, int half, int off, const float (&v)[4]) {
    float4 h, l;
    split_tf32(v[0], h.x, l.x); split_tf32(v[1], h.y, l.y); split_tf32(v[2], h.z, l.z); split_tf32(v[3], h.w, l.w);
    *reinterpret_cast<float4*>(tile + off) = h;
    *reinterpret_cast<float4*>(tile + half + off) = l;
}

// cotangent of one l=0 output channel m (gate VJP fused), times its norm
__device__ __forceinline__ float h_z(const TcBwdArgs& A, const float* rawr, const float* gr, const int* oz, const int* ov,
                                     const float* nz, int m) {
    if (m >= A.mz) return 0.0f;
    const int rc = oz[m];
    float h;
    if (A.epi.mode != SE3_EPI_GATE) h = gr[rc];
    else if (m < A.epi.ns_g) {
        const float x = rawr[rc], s = sigm(x);
        h = gr[m] * A.epi.cs * s * (1.0f + x * (1.0f - s));
    } else {
        const int v = m - A.epi.ns_g, vc = ov[v];
        const float s = sigm(rawr[rc]);
        const float* gg = gr + A.epi.ns_g + 3 * v;
        h = A.epi.cg * s * (1.0f - s) * (gg[0] * rawr[vc] + gg[1] * rawr[vc + 1] + gg[2] * rawr[vc + 2]);
    }
    return h * nz[m];
}
// cotangent of the three components of l=1 output channel v
__device__ __forceinline__ void h_v(const TcBwdArgs& A, const float* rawr, const float* gr, const int* oz, const int* ov,
                                    const float* nv, int v, float& a, float& b, float& c) {
    a = b = c = 0.0f;
    if (v >= A.mv) return;
    if (A.epi.mode != SE3_EPI_GATE) {
        const float* gg = gr + ov[v];
        a = gg[0]; b = gg[1]; c = gg[2];
    } else {
        const float s = A.epi.cg * sigm(rawr[oz[A.epi.ns_g + v]]);
        const float* gg = gr + A.epi.ns_g + 3 * v;
        a = s * gg[0]; b = s * gg[1]; c = s * gg[2];
    }
    a *= nv[3 * v]; b *= nv[3 * v + 1]; c *= nv[3 * v + 2];
}

__device__ __forceinline__ void build_h_task(const TcBwdArgs& A, unsigned char* smraw, const float* st, const int* tab,
                                             const float* norm, int rb, int gp, int lane) {
    const int row = rb * 16 + (lane & 15);
    const int g = gp * 2 + (lane >> 4);
    const int nzg = A.N2 >> 2, nvg = A.N3 >> 2;
    const bool gate = A.epi.mode == SE3_EPI_GATE;
    const int nsg = A.epi.ns_g;
    const float* rawr = st + A.rawoff + row * A.rawstride;
    const float* gr = st + A.goff + row * A.gstride;
    const float4 y = *reinterpret_cast<const float4*>(st + A.in2off + row * 4);
    const float* nz = norm;
    const float* nv = norm + A.mz;
    const int* oz = tab + A.t_oz;
    const int* ov = tab + A.t_ov;
    if (g < nzg) {
        float hz[4], hy[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = 4 * g + j;
            float h = 0.0f;
            if (m < A.mz) {
                const int rc = oz[m];
                if (!gate) h = gr[rc];
                else if (m < nsg) {
                    const float x = rawr[rc], s = sigm(x);
                    h = gr[m] * A.epi.cs * s * (1.0f + x * (1.0f - s));
                } else {
                    const int v = m - nsg, vc = ov[v];
                    const float s = sigm(rawr[rc]);
                    const float* gg = gr + nsg + 3 * v;
                    h = A.epi.cg * s * (1.0f - s) * (gg[0] * rawr[vc] + gg[1] * rawr[vc + 1] + gg[2] * rawr[vc + 2]);
                }
                h *= nz[m];
            }
            hz[j] = h;
            hy[j] = y.x * h;
        }
        const int off2 = (row >> 3) * A.rg2 + (g << 7) + ((row & 7) << 4);
        const int off1 = (row >> 3) * A.rg1 + (g << 7) + ((row & 7) << 4);
        put4(smraw + A.o_t2, A.sz_t2, off2, hz);
        put4(smraw + A.o_t1, A.sz_t1, off1, hy);
    } else if (g < nzg + nvg) {
        const int gv = g - nzg;
        float h0[4], h1[4], h2[4], hg[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int v = 4 * gv + j;
            float a = 0.f, b = 0.f, c = 0.f;
            if (v < A.mv) {
                if (!gate) {
                    const float* gg = gr + ov[v];
                    a = gg[0]; b = gg[1]; c = gg[2];
                } else {
                    const float s = A.epi.cg * sigm(rawr[oz[nsg + v]]);
                    const float* gg = gr + nsg + 3 * v;
                    a = s * gg[0]; b = s * gg[1]; c = s * gg[2];
                }
                a *= nv[3 * v]; b *= nv[3 * v + 1]; c *= nv[3 * v + 2];
            }
            h0[j] = a; h1[j] = b; h2[j] = c;
            hg[j] = C3f * (y.y * a + y.z * b + y.w * c);
        }
        const int off3 = (row >> 3) * A.rg3 + (gv << 7) + ((row & 7) << 4);
        put4(smraw + A.o_t3, A.sz_t3, off3, h0);
        put4(smraw + A.o_t3 + 2 * A.sz_t3, A.sz_t3, off3, h1);
        put4(smraw + A.o_t3 + 4 * A.sz_t3, A.sz_t3, off3, h2);
        const int off1 = (row >> 3) * A.rg1 + ((nzg + gv) << 7) + ((row & 7) << 4);
        put4(smraw + A.o_t1, A.sz_t1, off1, hg);
    }
}

// producer: one row per lane; segments (if with_x), in2, raw (gate) and the cotangent row
template <int ROWS, int NSLOT = 2>
__device__ __forceinline__ void producer_loop(const TcBwdArgs& A, unsigned char* smraw, uint32_t bar_full0,
                                              uint32_t bar_empty0, int prow, bool with_x, long long ntiles,
                                              int part = 0, int nparts = 1) {
    const long long R = A.rows;
    // row indices are fetched TWO tiles ahead (cur: next tile to load, nxt: the one after): with an almost always
    // empty stage the one-ahead prefetch was consumed immediately and its latency serialised with the row loads
    // (profiles/r01_v16_bwdw: the producers spent 75 % of their time waiting for index values)
    long long cur[SE3_MAX_SEG], nxt[SE3_MAX_SEG];
    long long curg, nxtg;
    auto fetch = [&](long long tile, long long (&c)[SE3_MAX_SEG], long long& cg) {
        const long long gr = tile * ROWS + prow;
#pragma unroll
        for (int s = 0; s < SE3_MAX_SEG; ++s)
            c[s] = (with_x && s < A.src.nseg && gr < R) ? (A.src.idx[s] ? (long long)A.src.idx[s][gr] : gr) : 0;
        cg = gr < R ? (A.gout_idx ? (long long)A.gout_idx[gr] : gr) : 0;
    };
    fetch(blockIdx.x, cur, curg);
    fetch((long long)blockIdx.x + gridDim.x, nxt, nxtg);
    int it = 0, slot = 0, use = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const long long gr = tile * ROWS + prow;
        const bool valid = gr < R;
        mbar_wait(bar_empty0 + 8 * slot, (use & 1) ^ 1);
        const uint32_t sbase = smem_u32(smraw) + A.o_stage + (uint32_t)slot * A.slot_floats * 4;
        if (with_x) {
#pragma unroll
            for (int s = 0; s < SE3_MAX_SEG; ++s) {
                if (s >= A.src.nseg) break;
                const float* srcp = A.src.base[s] + (valid ? cur[s] * A.src.ld[s] : 0);
                const uint32_t dst = sbase + (A.soff[s] + prow * A.sstride[s]) * 4;
                const int w = A.swidth[s];
                if (A.vec16[s]) for (int c = 4 * part; c < w; c += 4 * nparts) cp_async16(dst + c * 4, srcp + c, valid);
                else for (int c = part; c < w; c += nparts) cp_async4(dst + c * 4, srcp + c, valid);
            }
        }
        if (part == 0) cp_async16(sbase + (A.in2off + prow * 4) * 4, A.in2 + (valid ? gr * 4 : 0), valid);
        if (A.epi.mode == SE3_EPI_GATE) {
            const float* srcp = A.raw + (valid ? gr * A.d_out : 0);
            const uint32_t dst = sbase + (A.rawoff + prow * A.rawstride) * 4;
            if (A.graw_vec16) for (int c = 4 * part; c < A.d_out; c += 4 * nparts) cp_async16(dst + c * 4, srcp + c, valid);
            else for (int c = 2 * part; c < A.d_out; c += 2 * nparts) {
                const int sz = valid ? 8 : 0;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst + c * 4), "l"(srcp + c), "r"(sz) : "memory");
            }
        }
        {
            const float* srcp = A.gout + (valid ? curg * A.gwidth : 0);
            const uint32_t dst = sbase + (A.goff + prow * A.gstride) * 4;
            if (A.g_vec16) for (int c = 4 * part; c < A.gwidth; c += 4 * nparts) cp_async16(dst + c * 4, srcp + c, valid);
            else for (int c = part; c < A.gwidth; c += nparts) cp_async4(dst + c * 4, srcp + c, valid);
        }
        cp_async_mbar_arrive_noinc(bar_full0 + 8 * slot);
#pragma unroll
        for (int s = 0; s < SE3_MAX_SEG; ++s) cur[s] = nxt[s];
        curg = nxtg;
        fetch(tile + 2ll * gridDim.x, nxt, nxtg);
        if (++slot == NSLOT) { slot = 0; ++use; }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
}

// TMA producer (weight-gradient kernel): the stage is filled by cp.async.bulk copies that complete on the slot's
// mbarrier (complete_tx): one 256-byte copy per gathered row and segment, one contiguous copy per tile for in2, the
// saved pre-activation and narrow identity segments.  The LDGSTS producer above needs ~90 warp instructions and ~1100
// 32-byte sector requests per 32-row tile and was bound by outstanding requests, not by HBM (profiles/r01_final_*bwdw*).
// part 0: in1 segments; part 1: in2, pre-activation, cotangent.  The last, partial tile is copied by hand (zero fill).
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(bar), "r"(bytes) : "memory");
}
template <int ROWS, int NSLOT>
__device__ __forceinline__ void producer_loop_tma(const TcBwdArgs& A, unsigned char* smraw, uint32_t bar_full0,
                                                  uint32_t bar_empty0, int lane, long long ntiles, int part) {
    const long long R = A.rows;
    const bool gate = A.epi.mode == SE3_EPI_GATE;
    long long cur[SE3_MAX_SEG], nxt[SE3_MAX_SEG];
    long long curg, nxtg;
    auto fetch = [&](long long tile, long long (&c)[SE3_MAX_SEG], long long& cg) {
        long long gr = tile * ROWS + lane;
        if (gr > R - 1) gr = R - 1;
        if (gr < 0) gr = 0;
#pragma unroll
        for (int s = 0; s < SE3_MAX_SEG; ++s)
            c[s] = (part == 0 && s < A.src.nseg && tile < ntiles) ? (A.src.idx[s] ? (long long)A.src.idx[s][gr] : gr) : 0;
        cg = (part == 1 && tile < ntiles) ? (A.gout_idx ? (long long)A.gout_idx[gr] : gr) : 0;
    };
    fetch(blockIdx.x, cur, curg);
    fetch((long long)blockIdx.x + gridDim.x, nxt, nxtg);
    uint32_t bytes = 0;
    if (part == 0) {
        for (int s = 0; s < A.src.nseg; ++s) bytes += (uint32_t)(ROWS * A.swidth[s] * 4);
    } else {
        bytes = (uint32_t)(ROWS * 16 + (gate ? ROWS * A.d_out * 4 : 0) + ROWS * A.gwidth * 4);
    }
    int slot = 0, use = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        mbar_wait(bar_empty0 + 8 * slot, (use & 1) ^ 1);
        const uint32_t bar = bar_full0 + 8 * slot;
        unsigned char* sptr = smraw + A.o_stage + (size_t)slot * A.slot_floats * 4;
        const uint32_t sbase = smem_u32(sptr);
        const long long row0 = tile * ROWS;
        if (row0 + ROWS <= R) {
            if (lane == 0) mbar_arrive_tx(bar, bytes);
            __syncwarp();
            if (part == 0) {
#pragma unroll
                for (int s = 0; s < SE3_MAX_SEG; ++s) {
                    if (s >= A.src.nseg) break;
                    const int w = A.swidth[s];
                    if (A.vec16[s])
                        bulk_g2s(sbase + (A.soff[s] + lane * A.sstride[s]) * 4, A.src.base[s] + cur[s] * A.src.ld[s], (uint32_t)(w * 4), bar);
                    else if (lane == 0)
                        bulk_g2s(sbase + A.soff[s] * 4, A.src.base[s] + row0 * w, (uint32_t)(ROWS * w * 4), bar);
                }
            } else {
                if (lane == 0) bulk_g2s(sbase + A.in2off * 4, A.in2 + row0 * 4, (uint32_t)(ROWS * 16), bar);
                if (lane == 1 && gate) bulk_g2s(sbase + A.rawoff * 4, A.raw + row0 * A.d_out, (uint32_t)(ROWS * A.d_out * 4), bar);
                bulk_g2s(sbase + (A.goff + lane * A.gstride) * 4, A.gout + curg * A.gwidth, (uint32_t)(A.gwidth * 4), bar);
            }
        } else {
            // partial tile: plain loads / stores, zero rows past the end
            float* sf = reinterpret_cast<float*>(sptr);
            const bool valid = row0 + lane < R;
            if (part == 0) {
                for (int s = 0; s < A.src.nseg; ++s) {
                    const int w = A.swidth[s];
                    const float* srcp = A.src.base[s] + (valid ? cur[s] * A.src.ld[s] : 0);
                    float* d = sf + A.soff[s] + lane * A.sstride[s];
                    for (int c = 0; c < w; ++c) d[c] = valid ? __ldg(srcp + c) : 0.0f;
                }
            } else {
                const long long gr = row0 + lane;
                for (int c = 0; c < 4; ++c) sf[A.in2off + lane * 4 + c] = valid ? __ldg(A.in2 + gr * 4 + c) : 0.0f;
                if (gate)
                    for (int c = 0; c < A.d_out; ++c) sf[A.rawoff + lane * A.rawstride + c] = valid ? __ldg(A.raw + gr * A.d_out + c) : 0.0f;
                for (int c = 0; c < A.gwidth; ++c) sf[A.goff + lane * A.gstride + c] = valid ? __ldg(A.gout + curg * A.gwidth + c) : 0.0f;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar);
        }
#pragma unroll
        for (int s = 0; s < SE3_MAX_SEG; ++s) cur[s] = nxt[s];
        curg = nxtg;
        fetch(tile + 2ll * gridDim.x, nxt, nxtg);
        if (++slot == NSLOT) { slot = 0; ++use; }
    }
}

// tables shared by both kernels: plan tables, norms, packed (stride<<20 | offset) stage addresses of every scalar /
// vector channel
__device__ __forceinline__ void setup_tables(const TcBwdArgs& A, unsigned char* smraw, int nthreads) {
    int* tab = reinterpret_cast<int*>(smraw + A.o_tab);
    float* norm = reinterpret_cast<float*>(smraw + A.o_norm);
    int* stab = reinterpret_cast<int*>(smraw + A.o_tbl);
    int* vtab = stab + 8 * A.NSG8;
    const int tid = threadIdx.x;
    for (int t = tid; t < A.ntab; t += nthreads) tab[t] = A.tab[t];
    for (int t = tid; t < A.mz; t += nthreads) norm[t] = A.nz ? A.nz[t] : 1.0f;
    for (int t = tid; t < 3 * A.mv; t += nthreads) norm[A.mz + t] = A.nv ? A.nv[t] : 1.0f;
    __syncthreads();
    for (int k = tid; k < 8 * A.NSG8; k += nthreads) {
        int e = -1;
        if (k < A.ns) {
            const int vc = tab[A.t_s + k];
            int s = 0;
            for (int q = 1; q < SE3_MAX_SEG; ++q)
                if (q < A.src.nseg && vc >= A.src.cum[q]) s = q;
            e = (A.sstride[s] << 20) | (A.soff[s] + vc - A.src.cum[s]);
        }
        stab[k] = e;
    }
    for (int k = tid; k < 8 * A.NDG8; k += nthreads) {
        int e = -1;
        if (k < A.nd) {
            const int vc = tab[A.t_d + k];
            int s = 0;
            for (int q = 1; q < SE3_MAX_SEG; ++q)
                if (q < A.src.nseg && vc >= A.src.cum[q]) s = q;
            e = (A.sstride[s] << 20) | (A.soff[s] + vc - A.src.cum[s]);
        }
        vtab[k] = e;
    }
}

// ===================================================================== (W) weight-gradient kernel
__global__ void __launch_bounds__(TBW_THREADS, 1) l1tp_tc_bwdw_kernel(const TcBwdArgs A) {
    extern __shared__ __align__(1024) unsigned char smraw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* smf = reinterpret_cast<float*>(smraw);
    const int* tab = reinterpret_cast<const int*>(smraw + A.o_tab);
    const float* norm = reinterpret_cast<const float*>(smraw + A.o_norm);
    const int* stab = reinterpret_cast<const int*>(smraw + A.o_tbl);
    const int* vtab = stab + 8 * A.NSG8;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + A.o_bar);
    const uint32_t bar0 = smem_u32(bars);
    // barriers: 9..11 stage full | 12..14 stage empty (3-slot stage ring) | 4,7 half-set full | 5,8 half-set empty | 6 acc full.
    // The 32-row tile set is handed over in two 16-row halves (2 K-steps each), so the builders fill one half while the
    // tensor pipe consumes the other (profiles/r01_v8: 65 % of the builders' time was spent waiting for whole-set MMAs).
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);
    if (tid == 0) {
        for (int i = 0; i < BW_NSLOT; ++i) {
            mbar_init(BAR(9 + i), A.use_tma ? 2 : 64);  // TMA: one expect_tx arrival per producer warp; else 64 producer lanes
            mbar_init(BAR(12 + i), BW_NBUILD);
        }
        mbar_init(BAR(4), BW_NBUILD);
        mbar_init(BAR(5), 1);
        mbar_init(BAR(6), 1);
        mbar_init(BAR(7), BW_NBUILD);
        mbar_init(BAR(8), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    setup_tables(A, smraw, TBW_THREADS);
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(A.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long long R = A.rows;
    const long long ntiles = (R + TMB - 1) / TMB;
    const int nt_cta = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    // TMEM columns
    const int cD1 = 0, cD2 = A.N2, cD3 = A.N2 + A.N3, cD4 = 2 * A.N2 + A.N3;

    if (warp == 0 || warp == 2) {
        if (A.use_tma) producer_loop_tma<TMB, BW_NSLOT>(A, smraw, BAR(9), BAR(12), lane, ntiles, warp == 0 ? 0 : 1);
        else producer_loop<TMB, BW_NSLOT>(A, smraw, BAR(9), BAR(12), lane, true, ntiles, warp == 0 ? 0 : 1, 2);
    } else if (warp == 1) {
        // ---------------- MMA issuer: per tile 4 K-steps (8 rows each) x {S.[HZY|HG], Dd.HZ, AVc.HVc} x 3xTF32.
        // All operands are K-major with K = rows: transposed tiles [channel/feature][row], 8-channel groups of
        // GS = (TMB/4)*128 bytes, the two 16-byte K-chunks of one MMA 128 bytes apart.  The tile set is single
        // buffered, so every descriptor is tile-invariant: they are built once and the loop only adds constants
        // (profiles/r01_v8: descriptor arithmetic made the issue loop 85 cycles per MMA, 3x the tensor-pipe floor).
        const uint32_t sb = smem_u32(smraw);
        const uint32_t GS = (TMB / 4) * 128;
        // S.[Y0 HZ | HG] is one MMA: the HG groups follow the HZ groups in the T1 tile and cD2 = cD1 + N2
        const uint32_t idS = make_idesc_ex(A.MS, A.N2 + A.N3, 0, 0);
        const uint32_t idD = make_idesc_ex(64, A.N2, 0, 0), idV = make_idesc_ex(64, A.N3, 0, 0);
        const uint64_t dSh = make_desc(sb + A.o_as, GS), dSl = make_desc(sb + A.o_as + A.sz_as, GS);
        const uint64_t dDh = make_desc(sb + A.o_ad, GS), dDl = make_desc(sb + A.o_ad + A.sz_ad, GS);
        const uint64_t dVh = make_desc(sb + A.o_av, GS), dVl = make_desc(sb + A.o_av + A.sz_ad, GS);
        const uint64_t d1h = make_desc(sb + A.o_t1, GS), d1l = make_desc(sb + A.o_t1 + A.sz_t1, GS);
        const uint64_t d2h = make_desc(sb + A.o_t2, GS), d2l = make_desc(sb + A.o_t2 + A.sz_t2, GS);
        const uint64_t d3h = make_desc(sb + A.o_t3, GS), d3l = make_desc(sb + A.o_t3 + A.sz_t3, GS);
        const uint64_t vA = (uint64_t)((2u * A.sz_ad) >> 4), vB = (uint64_t)((2u * A.sz_t3) >> 4);
        for (int it = 0; it < 2 * nt_cta; ++it) {
            const int hf = it & 1, tl = it >> 1;
            mbar_wait(BAR(hf ? 7 : 4), tl & 1);
            tc_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int k2 = 0; k2 < TMB / 16; ++k2) {
                    const int ks = hf * (TMB / 16) + k2;
                    const uint32_t acc0 = (it == 0 && k2 == 0) ? 0u : 1u;
                    const uint64_t ko = (uint64_t)(ks * 16);   // 256 bytes per K-step in 16-byte units
                    tc_mma_tf32(tmem_base + cD1, dSh + ko, d1h + ko, idS, acc0);
                    tc_mma_tf32(tmem_base + cD1, dSh + ko, d1l + ko, idS, 1u);
                    tc_mma_tf32(tmem_base + cD1, dSl + ko, d1h + ko, idS, 1u);
                    tc_mma_tf32(tmem_base + cD3, dDh + ko, d2h + ko, idD, acc0);
                    tc_mma_tf32(tmem_base + cD3, dDh + ko, d2l + ko, idD, 1u);
                    tc_mma_tf32(tmem_base + cD3, dDl + ko, d2h + ko, idD, 1u);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const uint64_t oa = ko + (uint64_t)c * vA, ob = ko + (uint64_t)c * vB;
                        tc_mma_tf32(tmem_base + cD4, dVh + oa, d3h + ob, idV, c == 0 ? acc0 : 1u);
                        tc_mma_tf32(tmem_base + cD4, dVh + oa, d3l + ob, idV, 1u);
                        tc_mma_tf32(tmem_base + cD4, dVl + oa, d3h + ob, idV, 1u);
                    }
                }
                tc_commit(BAR(hf ? 8 : 5));
            }
            __syncwarp();
        }
        if (lane == 0) tc_commit(BAR(6));
        __syncwarp();
    } else if (warp >= BW_BUILD_W0) {
        // ---------------- builders: transposed feature (S, Dd, AVc) and cotangent (T1, T2, T3c) tiles for 32 rows.
        // warp-task = one 8-channel group x 16 rows: lane = (channel f8 = lane & 7, row quad q = lane >> 3); a thread
        // gathers its channel for 4 consecutive rows and writes one 16-byte K-chunk (hi and lo).
        const int bw = warp - BW_BUILD_W0;
        const int f8 = lane & 7, q = lane >> 3;
        const int nrb = TMB / 16;
        const int nzg8 = A.N2 >> 3, nvg8 = A.N3 >> 3;
        const int ntask = nrb * (A.NSG8 + A.NDG8 + nzg8 + nvg8);
        const int GS = (TMB / 4) * 128;
        const int* oz = tab + A.t_oz;
        const int* ov = tab + A.t_ov;
        const float* nz = norm;
        const float* nv = norm + A.mz;
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int slot = it % BW_NSLOT, use = it / BW_NSLOT;
            mbar_wait(BAR(9 + slot), use & 1);              // stage full
            const float* st = smf + (A.o_stage >> 2) + (size_t)slot * A.slot_floats;
            for (int rb = 0; rb < nrb; ++rb) {
            mbar_wait(BAR(rb ? 8 : 5), (it & 1) ^ 1);       // half-set free (its MMAs of the previous tile are done)
            for (int t = bw; t < ntask / nrb; t += BW_NBUILD) {
                int g = t;
                const int r0 = rb * 16 + q * 4;                       // first of this thread's 4 rows
                const int off = g * GS + ((rb * 4 + q) << 7) + (f8 << 4);  // (group, K-chunk, channel) -- g rebased below
                if (g < A.NSG8) {
                    const int ee = stab[8 * g + f8];
                    float v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i] = ee >= 0 ? st[(ee & 0xfffff) + (r0 + i) * (ee >> 20)] : 0.0f;
                    put4(smraw + A.o_as, A.sz_as, off, v);
                } else if (g < A.NSG8 + A.NDG8) {
                    g -= A.NSG8;
                    const int o2 = g * GS + ((rb * 4 + q) << 7) + (f8 << 4);
                    const int ee = vtab[8 * g + f8];
                    float d[4], a0[4], a1[4], a2[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 y = *reinterpret_cast<const float4*>(st + A.in2off + (r0 + i) * 4);
                        float vx = 0.f, vy = 0.f, vz = 0.f;
                        if (ee >= 0) {
                            const float* p = st + (ee & 0xfffff) + (r0 + i) * (ee >> 20);
                            vx = p[0]; vy = p[1]; vz = p[2];
                        }
                        const float sy0 = C3f * y.x;
                        d[i] = C3f * (vx * y.y + vy * y.z + vz * y.w);
                        a0[i] = sy0 * vx; a1[i] = sy0 * vy; a2[i] = sy0 * vz;
                    }
                    put4(smraw + A.o_ad, A.sz_ad, o2, d);
                    put4(smraw + A.o_av, A.sz_ad, o2, a0);
                    put4(smraw + A.o_av + 2 * A.sz_ad, A.sz_ad, o2, a1);
                    put4(smraw + A.o_av + 4 * A.sz_ad, A.sz_ad, o2, a2);
                } else if (g < A.NSG8 + A.NDG8 + nzg8) {
                    g -= A.NSG8 + A.NDG8;
                    const int o2 = g * GS + ((rb * 4 + q) << 7) + (f8 << 4);
                    const int m = 8 * g + f8;
                    float hz[4], hy[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int row = r0 + i;
                        hz[i] = h_z(A, st + A.rawoff + row * A.rawstride, st + A.goff + row * A.gstride, oz, ov, nz, m);
                        hy[i] = st[A.in2off + row * 4] * hz[i];
                    }
                    put4(smraw + A.o_t2, A.sz_t2, o2, hz);
                    put4(smraw + A.o_t1, A.sz_t1, o2, hy);
                } else {
                    g -= A.NSG8 + A.NDG8 + nzg8;
                    const int o2 = g * GS + ((rb * 4 + q) << 7) + (f8 << 4);
                    const int v = 8 * g + f8;
                    float h0[4], h1[4], h2[4], hg[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int row = r0 + i;
                        const float4 y = *reinterpret_cast<const float4*>(st + A.in2off + row * 4);
                        h_v(A, st + A.rawoff + row * A.rawstride, st + A.goff + row * A.gstride, oz, ov, nv, v, h0[i], h1[i], h2[i]);
                        hg[i] = C3f * (y.y * h0[i] + y.z * h1[i] + y.w * h2[i]);
                    }
                    put4(smraw + A.o_t3, A.sz_t3, o2, h0);
                    put4(smraw + A.o_t3 + 2 * A.sz_t3, A.sz_t3, o2, h1);
                    put4(smraw + A.o_t3 + 4 * A.sz_t3, A.sz_t3, o2, h2);
                    put4(smraw + A.o_t1, A.sz_t1, (nzg8 + g) * GS + ((rb * 4 + q) << 7) + (f8 << 4), hg);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(rb ? 7 : 4));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(12 + slot));
        }
        // ---------------- final epilogue (warps 4-7): TMEM accumulators -> per-CTA partials
        if (bw < 4) {
            const int e = warp & 3;
            mbar_wait(BAR(6), 0);
            tc_fence_after();
            float* part = A.partials + (long long)blockIdx.x * A.wtot;
            const uint32_t tq = tmem_base + ((uint32_t)(32 * e) << 16);
            // rows of the S operand
            const int ks = A.MS == 128 ? 32 * e + lane : 16 * e + (lane & 15);
            const bool ks_ok = (A.MS == 128 || lane < 16) && ks < A.ns;
            const int kd = 16 * e + (lane & 15);
            const bool kd_ok = lane < 16 && kd < A.nd;
            for (int m0 = 0; m0 < A.N2; m0 += 8) {
                float a[8], b[8];
                tc_ld8(tq + cD1 + m0, a);
                tc_ld8(tq + cD3 + m0, b);
                tc_wait_ld();
#ifdef SE3_TC_DEBUG
                if (blockIdx.x == 0 && m0 == 0 && (lane < 3 || lane == 16))
                    printf("dbg e=%d lane=%d D1: %g %g %g %g | D3: %g %g %g %g | A_S[0..3]=%g %g %g %g T1=%g %g T2=%g %g\n", e, lane, a[0], a[1], a[2],
                           a[3], b[0], b[1], b[2], b[3], ((float*)(smraw + A.o_as))[0], ((float*)(smraw + A.o_as))[1],
                           ((float*)(smraw + A.o_as))[4], ((float*)(smraw + A.o_as))[5], ((float*)(smraw + A.o_t1))[0],
                           ((float*)(smraw + A.o_t1))[1], ((float*)(smraw + A.o_t2))[0], ((float*)(smraw + A.o_t2))[1]);
#endif
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int m = m0 + j;
                    if (m < A.mz) {
                        if (ks_ok) part[A.gw_z_off + (long long)ks * A.mz + m] = a[j];
                        if (kd_ok) part[A.gw_z_off + (long long)(A.ns + kd) * A.mz + m] = b[j];
                    }
                }
            }
            for (int m0 = 0; m0 < A.N3; m0 += 8) {
                float a[8], b[8];
                tc_ld8(tq + cD2 + m0, a);
                tc_ld8(tq + cD4 + m0, b);
                tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int m = m0 + j;
                    if (m < A.mv) {
                        if (ks_ok) part[A.gw_v_off + (long long)ks * A.mv + m] = a[j];
                        if (kd_ok) part[A.gw_v_off + (long long)(A.ns + kd) * A.mv + m] = b[j];
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(A.tmem_cols) : "memory");
    }
}


// ===================================================================== (I) input-gradient kernel
static constexpr int TBI_THREADS = 768;  // warps: 0,2 producers | 1 MMA | 4-15 builders | 16-23 epilogue
static constexpr int TMI = 64;
static constexpr int BI_BUILD_W0 = 4, BI_NBUILD = 12, BI_EPI_W0 = 16, BI_NEPI = 8;

__global__ void __launch_bounds__(TBI_THREADS, 1) l1tp_tc_bwdi_kernel(const TcBwdArgs A) {
    extern __shared__ __align__(1024) unsigned char smraw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* smf = reinterpret_cast<float*>(smraw);
    const int* tab = reinterpret_cast<const int*>(smraw + A.o_tab);
    const float* norm = reinterpret_cast<const float*>(smraw + A.o_norm);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + A.o_bar);
    const uint32_t bar0 = smem_u32(bars);
    // barriers: 0,1 stage full | 2,3 stage empty | 4 H full | 5 H empty | 6 acc full
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(BAR(i), TMI);
            mbar_init(BAR(2 + i), BI_NBUILD);
        }
        mbar_init(BAR(4), BI_NBUILD);
        mbar_init(BAR(5), BI_NEPI);
        mbar_init(BAR(6), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    setup_tables(A, smraw, TBI_THREADS);
    // W^T tiles (B operands, K-major with K = output channels), hi | lo
    {
        const int K1 = A.N2 + A.N3, NS = A.NS8 * 8, ND = A.ND8 * 8;
        unsigned char* b1 = smraw + A.o_b1;
        const int half1 = NS * K1 * 4;
        for (int t = tid; t < NS * K1; t += TBI_THREADS) {
            const int n = t / K1, ch = t - n * K1;
            float x = 0.0f;
            if (n < A.ns) {
                if (ch < A.mz) x = __ldg(A.wz + (long long)n * A.mz + ch);
                else if (ch >= A.N2 && ch - A.N2 < A.mv) x = __ldg(A.wv + (long long)n * A.mv + (ch - A.N2));
            }
            float hi, lo;
            split_tf32(x, hi, lo);
            const int o = canon_off(n, ch, K1 >> 2);
            *reinterpret_cast<float*>(b1 + o) = hi;
            *reinterpret_cast<float*>(b1 + half1 + o) = lo;
        }
        unsigned char* b2 = smraw + A.o_b2;
        const int half2 = ND * A.N2 * 4;
        for (int t = tid; t < ND * A.N2; t += TBI_THREADS) {
            const int n = t / A.N2, ch = t - n * A.N2;
            float x = 0.0f;
            if (n < A.nd && ch < A.mz) x = __ldg(A.wz + (long long)(A.ns + n) * A.mz + ch);
            float hi, lo;
            split_tf32(x, hi, lo);
            const int o = canon_off(n, ch, A.N2 >> 2);
            *reinterpret_cast<float*>(b2 + o) = hi;
            *reinterpret_cast<float*>(b2 + half2 + o) = lo;
        }
        unsigned char* b3 = smraw + A.o_b3;
        const int half3 = ND * A.N3 * 4;
        for (int t = tid; t < ND * A.N3; t += TBI_THREADS) {
            const int n = t / A.N3, ch = t - n * A.N3;
            float x = 0.0f;
            if (n < A.nd && ch < A.mv) x = __ldg(A.wv + (long long)(A.ns + n) * A.mv + ch);
            float hi, lo;
            split_tf32(x, hi, lo);
            const int o = canon_off(n, ch, A.N3 >> 2);
            *reinterpret_cast<float*>(b3 + o) = hi;
            *reinterpret_cast<float*>(b3 + half3 + o) = lo;
        }
    }
    fence_proxy_async();
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(A.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long long R = A.rows;
    const long long ntiles = (R + TMI - 1) / TMI;
    const int NS = A.NS8 * 8, ND = A.ND8 * 8;
    const int cS = 0, cD = NS, cT = NS + ND;  // TMEM columns: gS | gD | gT0 gT1 gT2

    if (warp == 0 || warp == 2) {
        producer_loop<TMI>(A, smraw, BAR(0), BAR(2), (warp == 0 ? 0 : 32) + lane, false, ntiles);
    } else if (warp == 1) {
        // descriptors are tile-invariant (single-buffered H tiles): build them once, add constants in the loop
        const uint32_t sb = smem_u32(smraw);
        const uint32_t id1 = make_idesc(NS), id2 = make_idesc(ND);
        const uint32_t K1 = A.N2 + A.N3;
        const uint32_t sbo_b1 = (K1 >> 2) * 128, sbo_b2 = (A.N2 >> 2) * 128, sbo_b3 = (A.N3 >> 2) * 128;
        const uint32_t hb1 = NS * K1 * 4, hb2 = ND * A.N2 * 4, hb3 = ND * A.N3 * 4;
        const uint64_t a1h = make_desc(sb + A.o_t1, A.rg1), a1l = make_desc(sb + A.o_t1 + A.sz_t1, A.rg1);
        const uint64_t a2h = make_desc(sb + A.o_t2, A.rg2), a2l = make_desc(sb + A.o_t2 + A.sz_t2, A.rg2);
        const uint64_t a3h = make_desc(sb + A.o_t3, A.rg3), a3l = make_desc(sb + A.o_t3 + A.sz_t3, A.rg3);
        const uint64_t b1h = make_desc(sb + A.o_b1, sbo_b1), b1l = make_desc(sb + A.o_b1 + hb1, sbo_b1);
        const uint64_t b2h = make_desc(sb + A.o_b2, sbo_b2), b2l = make_desc(sb + A.o_b2 + hb2, sbo_b2);
        const uint64_t b3h = make_desc(sb + A.o_b3, sbo_b3), b3l = make_desc(sb + A.o_b3 + hb3, sbo_b3);
        const uint64_t v3 = (uint64_t)((2u * A.sz_t3) >> 4);
        const int n1 = (int)(K1 >> 3), n2 = A.N2 >> 3, n3 = A.N3 >> 3;
        const int nt_cta = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
        for (int it = 0; it < nt_cta; ++it) {
            mbar_wait(BAR(4), it & 1);
            tc_fence_after();
            if (lane == 0) {
                for (int j = 0; j < n1; ++j) {
                    const uint64_t o = (uint64_t)(j * 16);
                    tc_mma_tf32(tmem_base + cS, a1h + o, b1h + o, id1, j ? 1u : 0u);
                    tc_mma_tf32(tmem_base + cS, a1h + o, b1l + o, id1, 1u);
                    tc_mma_tf32(tmem_base + cS, a1l + o, b1h + o, id1, 1u);
                }
                for (int j = 0; j < n2; ++j) {
                    const uint64_t o = (uint64_t)(j * 16);
                    tc_mma_tf32(tmem_base + cD, a2h + o, b2h + o, id2, j ? 1u : 0u);
                    tc_mma_tf32(tmem_base + cD, a2h + o, b2l + o, id2, 1u);
                    tc_mma_tf32(tmem_base + cD, a2l + o, b2h + o, id2, 1u);
                }
                for (int c = 0; c < 3; ++c)
                    for (int j = 0; j < n3; ++j) {
                        const uint64_t o = (uint64_t)(j * 16), oa = o + (uint64_t)c * v3;
                        tc_mma_tf32(tmem_base + cT + c * ND, a3h + oa, b3h + o, id2, j ? 1u : 0u);
                        tc_mma_tf32(tmem_base + cT + c * ND, a3h + oa, b3l + o, id2, 1u);
                        tc_mma_tf32(tmem_base + cT + c * ND, a3l + oa, b3h + o, id2, 1u);
                    }
                tc_commit(BAR(6));
            }
            __syncwarp();
        }
    } else if (warp >= BI_BUILD_W0 && warp < BI_BUILD_W0 + BI_NBUILD) {
        const int bw = warp - BI_BUILD_W0;
        const int ngp = ((A.N2 >> 2) + (A.N3 >> 2) + 1) >> 1;
        const int ntask = (TMI / 16) * ngp;
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int slot = it & 1, use = it >> 1;
            mbar_wait(BAR(slot), use & 1);       // stage full
            mbar_wait(BAR(5), (it & 1) ^ 1);     // H tiles / g tile free (finish of the previous tile done)
            const float* st = smf + (A.o_stage >> 2) + (size_t)slot * A.slot_floats;
            for (int t = bw; t < ntask; t += BI_NBUILD) build_h_task(A, smraw, st, tab, norm, t % (TMI / 16), t / (TMI / 16), lane);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(BAR(4));
                mbar_arrive(BAR(2 + slot));
            }
        }
    } else if (warp >= BI_EPI_W0) {
        const int e = warp & 3, h = (warp - BI_EPI_W0) >> 2;
        const int et = tid - BI_EPI_W0 * 32;
        const bool rowlane = lane < 16;
        const int row = 16 * e + (lane & 15);
        float* gt = reinterpret_cast<float*>(smraw + A.o_gt);
        const int gts = A.gts;
        const int* scol = tab + A.t_s;
        const int* vcol = tab + A.t_d;
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const long long row0 = tile * TMI;
            const int nvalid = (int)min((long long)TMI, R - row0);
            const long long gr = row0 + row;
            float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rowlane && gr < R) y = __ldg(reinterpret_cast<const float4*>(A.in2) + gr);
            mbar_wait(BAR(6), it & 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + ((uint32_t)(32 * e) << 16);
            float* grow = gt + row * gts;
            if (h == 0) {
                for (int k0 = 0; k0 < NS; k0 += 8) {
                    float v[8];
                    tc_ld8(acc + cS + k0, v);
                    tc_wait_ld();
                    if (rowlane) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (k0 + j < A.ns) grow[scol[k0 + j]] = v[j];
                    }
                }
            } else {
                const float sy0 = C3f * y.x, s1 = C3f * y.y, s2 = C3f * y.z, s3 = C3f * y.w;
                for (int k0 = 0; k0 < ND; k0 += 8) {
                    float d[8], t0[8], t1[8], t2[8];
                    tc_ld8(acc + cD + k0, d);
                    tc_ld8(acc + cT + k0, t0);
                    tc_ld8(acc + cT + ND + k0, t1);
                    tc_ld8(acc + cT + 2 * ND + k0, t2);
                    tc_wait_ld();
                    if (rowlane) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (k0 + j < A.nd) {
                                float* o = grow + vcol[k0 + j];
                                o[0] = fmaf(s1, d[j], sy0 * t0[j]);
                                o[1] = fmaf(s2, d[j], sy0 * t1[j]);
                                o[2] = fmaf(s3, d[j], sy0 * t2[j]);
                            }
                    }
                }
            }
            tc_fence_before();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // ---- scatter the input-gradient tile per segment
            for (int s = 0; s < A.src.nseg; ++s) {
                float* gb = A.gseg[s];
                const int mode = A.gmode[s];
                if (!gb || mode == SE3_GRAD_NONE) continue;
                const int w = A.src.cum[s + 1] - A.src.cum[s], c0 = A.src.cum[s], ld = A.src.ld[s];
                const int32_t* idx = A.src.idx[s];
                const bool v4 = (w & 3) == 0 && (ld & 3) == 0 && (c0 & 3) == 0 && ((uintptr_t)gb & 15) == 0;
                if (mode == SE3_GRAD_STORE || mode == SE3_GRAD_ATOMIC) {
                    if (v4) {
                        const int w4 = w >> 2;
                        for (int t = et; t < nvalid * w4; t += 256) {
                            const int r = t / w4, c = (t - r * w4) << 2;
                            const float4 v = *reinterpret_cast<const float4*>(gt + r * gts + c0 + c);
                            const long long dr = idx ? (long long)__ldg(idx + row0 + r) : row0 + r;
                            float* dst = gb + dr * ld + c;
                            if (mode == SE3_GRAD_STORE) *reinterpret_cast<float4*>(dst) = v;
                            else red_add_v4(dst, v.x, v.y, v.z, v.w);
                        }
                    } else {
                        for (int t = et; t < nvalid * w; t += 256) {
                            const int r = t / w, c = t - r * w;
                            const float v = gt[r * gts + c0 + c];
                            const long long dr = idx ? (long long)__ldg(idx + row0 + r) : row0 + r;
                            if (mode == SE3_GRAD_STORE) gb[dr * ld + c] = v;
                            else atomicAdd(gb + dr * ld + c, v);
                        }
                    }
                } else {  // SORTED: run-length combine equal destinations, one red per run
                    int parts = 256 / w;
                    if (parts < 1) parts = 1;
                    const int rpp = (TMI + parts - 1) / parts;
                    for (int item = et; item < w * parts; item += 256) {
                        const int c = item % w, qd = item / w;
                        const int rbeg = qd * rpp;
                        const int rend = min(rbeg + rpp, nvalid);
                        if (rbeg >= rend) continue;
                        int cur = __ldg(idx + row0 + rbeg);
                        float accv = 0.0f;
                        for (int r = rbeg; r < rend; ++r) {
                            const int k = __ldg(idx + row0 + r);
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
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(5));
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(A.tmem_cols) : "memory");
    }
}

}  // namespace se3

using namespace se3;

static int fill_common(TcBwdArgs& A, const int n[4], const int m[4], const int t_in[4], const int t_out[4], int ntab,
                       const int* d_tab, const se3_l1tp_bwd_args* a, const RowSrc& src, const EpiL& epi, int rows_per_tile,
                       bool with_x, bool want_tma = false) {
    if (n[1] || n[2] || m[1] || m[2]) return 1;
    const int ns = n[0], nd = n[3], mz = m[0], mv = m[3];
    if (ns < 1 || nd < 1 || mz < 1 || mv < 1) return 1;
    memset(&A, 0, sizeof(A));
    A.ns = ns; A.nd = nd; A.mz = mz; A.mv = mv;
    A.NSG = (ns + 3) >> 2; A.NDG = (nd + 3) >> 2;
    A.NSG8 = (ns + 7) >> 3; A.NDG8 = (nd + 7) >> 3;
    A.NS8 = A.NSG8; A.ND8 = A.NDG8;
    A.N2 = (mz + 7) & ~7; A.N3 = (mv + 7) & ~7;
    if (A.NSG8 > 16 || A.NDG8 > 8 || A.N2 + A.N3 > 256) return 1;
    A.MS = A.NSG8 > 8 ? 128 : 64;
    if (A.MS == 128 && ((A.N2 + A.N3) & 15)) return 1;   // M=128 needs N % 16 == 0 (S.[HZY|HG] is one MMA)
    A.rows = a->rows; A.src = src; A.in2 = a->in2; A.wz = a->w[0]; A.wv = a->w[3]; A.nz = a->norm[0]; A.nv = a->norm[3];
    A.epi = epi; A.raw = a->raw; A.gout = a->gout; A.gout_idx = a->gout_idx; A.tab = d_tab; A.ntab = ntab;
    A.t_s = t_in[0]; A.t_d = t_in[3]; A.t_oz = t_out[0]; A.t_ov = t_out[3];
    A.d_out = mz + 3 * mv;
    A.gwidth = epi.d_post;
    if (((uintptr_t)a->in2 & 15) != 0) return 1;
    // TMA (cp.async.bulk) staging: wide segments and the cotangent go row by row (256-byte rows), narrow identity
    // segments / in2 / the saved pre-activation as ONE contiguous copy per tile (their stage layout is then dense)
    bool tma = want_tma && with_x && ((epi.d_post & 3) == 0) && (((uintptr_t)a->gout & 15) == 0);
    if (tma && epi.mode == SE3_EPI_GATE && (((rows_per_tile * (mz + 3 * mv) * 4) & 15) || ((uintptr_t)a->raw & 15))) tma = false;
    for (int s = 0; s < src.nseg && tma; ++s) {
        const int w = src.cum[s + 1] - src.cum[s];
        const bool wide = (w & 3) == 0 && (src.ld[s] & 3) == 0 && ((uintptr_t)src.base[s] & 15) == 0;
        if (!wide && (src.idx[s] || src.ld[s] != w || ((rows_per_tile * w * 4) & 15) || ((uintptr_t)src.base[s] & 15))) tma = false;
    }
    {   // measured: with two or more GATHERED wide segments (msg1: x[dst], x[src]) the 256-byte per-row bulk copies are no
        // faster than the LDGSTS producer (1.25 vs 1.20 ms), with at most one they are (msg2: 0.93 vs 1.08 ms)
        int gathered = 0;
        for (int s = 0; s < src.nseg; ++s) gathered += (src.idx[s] != nullptr) ? 1 : 0;
        if (gathered >= 2) tma = false;
    }
    A.use_tma = tma ? 1 : 0;
    int off = 0;
    for (int s = 0; s < src.nseg; ++s) {
        const int w = src.cum[s + 1] - src.cum[s];
        A.swidth[s] = w;
        A.vec16[s] = ((w & 3) == 0 && (src.ld[s] & 3) == 0 && ((uintptr_t)src.base[s] & 15) == 0) ? 1 : 0;
        A.sstride[s] = (tma && !A.vec16[s]) ? w : tc_stage_stride(w);
        if (with_x) {
            if (!A.vec16[s] && w > 16) return 1;
            A.soff[s] = off;
            off += (rows_per_tile * A.sstride[s] + 3) & ~3;
        }
    }
    A.in2off = off; off += rows_per_tile * 4;
    A.rawstride = tma ? A.d_out : tc_stage_stride(A.d_out);
    A.rawoff = off;
    if (epi.mode == SE3_EPI_GATE) {
        if ((A.d_out & 1) || ((uintptr_t)a->raw & 7)) return 1;
        A.graw_vec16 = ((A.d_out & 3) == 0 && ((uintptr_t)a->raw & 15) == 0) ? 1 : 0;
        off += rows_per_tile * A.rawstride;
    }
    A.gstride = tc_stage_stride(A.gwidth);
    A.goff = off; off += rows_per_tile * A.gstride;
    A.g_vec16 = ((A.gwidth & 3) == 0 && ((uintptr_t)a->gout & 15) == 0) ? 1 : 0;
    if (!A.g_vec16 && A.gwidth > 16) return 1;
    A.slot_floats = off;
    return 0;
}

// weight gradients on the tensor cores; returns launched=false when the configuration is not eligible
int se3_l1tp_tc_try_backward_w(const int n[4], const int m[4], const int t_in[4], const int t_out[4], int ntab,
                               const int* d_tab, const se3_l1tp_bwd_args* a, const RowSrc& src, const EpiL& epi,
                               float* partials, int wtot, int gw_z_off, int gw_v_off, int max_grid, cudaStream_t st,
                               int* grid_out, bool* launched) {
    *launched = false;
    static int disabled = -1;
    if (disabled < 0) {
        const char* e = getenv("SE3_DISABLE_TC");
        disabled = (e && (e[0] == '1' || e[0] == '2')) ? 1 : 0;
    }
    if (disabled) return SE3_OK;
    TcBwdArgs A;
    static int no_tma = -1;
    if (no_tma < 0) { const char* e = getenv("SE3_DISABLE_TMA"); no_tma = (e && e[0] == '1') ? 1 : 0; }
    if (fill_common(A, n, m, t_in, t_out, ntab, d_tab, a, src, epi, TMB, true, !no_tma)) return SE3_OK;
    A.partials = partials; A.wtot = wtot; A.gw_z_off = gw_z_off; A.gw_v_off = gw_v_off;
    A.tmem_cols = 32;
    while (A.tmem_cols < 2 * A.N2 + 2 * A.N3) A.tmem_cols <<= 1;
    if (A.tmem_cols > 512) return SE3_OK;
    auto al = [](int x, int q) { return (x + q - 1) / q * q; };
    const int GS = (TMB / 4) * 128;  // bytes of one 8-channel group of a transposed tile
    A.sz_as = A.NSG8 * GS; A.sz_ad = A.NDG8 * GS;
    A.sz_t1 = ((A.N2 + A.N3) >> 3) * GS; A.sz_t2 = (A.N2 >> 3) * GS; A.sz_t3 = (A.N3 >> 3) * GS;
    int o = 0;
    A.o_as = o; o += 2 * A.sz_as;
    A.o_ad = o; o += 2 * A.sz_ad;
    A.o_av = o; o += 6 * A.sz_ad;
    A.o_t1 = o; o += 2 * A.sz_t1;
    A.o_t2 = o; o += 2 * A.sz_t2;
    A.o_t3 = o; o += 6 * A.sz_t3;
    o += 16 * GS;  // slack: M=64/128 operand reads run past the last real 8-channel group (those D rows are ignored)
    o = al(o, 128);
    A.o_stage = o; o += BW_NSLOT * A.slot_floats * 4;
    A.o_tab = o; o += al(ntab * 4, 16);
    A.o_norm = o; o += al((A.mz + 3 * A.mv) * 4, 16);
    A.o_tbl = o; o += al((8 * A.NSG8 + 8 * A.NDG8) * 4, 16);
    A.o_bar = o; o += 16 * 8 + 16;
    int dev = 0, maxsm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (o > maxsm) return SE3_OK;
    const int smem = std::max(o, 120 * 1024);
    static bool attr_set = false;
    if (!attr_set) {
        SE3_CUDA_TRY(cudaFuncSetAttribute(l1tp_tc_bwdw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
        attr_set = true;
    }
    const long long ntiles = (a->rows + TMB - 1) / TMB;
    const int grid = (int)std::min<long long>(ntiles, std::min(num_sms(), max_grid));
    if (getenv("SE3_DEBUG_NANFILL")) cudaMemsetAsync(partials, 0xFF, sizeof(float) * (size_t)grid * wtot, st);
    l1tp_tc_bwdw_kernel<<<grid, TBW_THREADS, smem, st>>>(A);
    SE3_LAUNCHED();
    g_tc_launches.fetch_add(1, std::memory_order_relaxed);
    *grid_out = grid;
    *launched = true;
    return SE3_OK;
}

// input gradients on the tensor cores; returns launched=false when the configuration is not eligible
int se3_l1tp_tc_try_backward_in(const int n[4], const int m[4], const int t_in[4], const int t_out[4], int ntab,
                                const int* d_tab, const se3_l1tp_bwd_args* a, const RowSrc& src, const EpiL& epi,
                                float* const gseg[SE3_MAX_SEG], const int gmode[SE3_MAX_SEG], cudaStream_t st,
                                bool* launched) {
    *launched = false;
    static int disabled = -1;
    if (disabled < 0) {
        const char* e = getenv("SE3_DISABLE_TC");
        disabled = (e && (e[0] == '1' || e[0] == '3')) ? 1 : 0;
    }
    if (disabled) return SE3_OK;
    TcBwdArgs A;
    if (fill_common(A, n, m, t_in, t_out, ntab, d_tab, a, src, epi, TMI, false)) return SE3_OK;
    for (int s = 0; s < SE3_MAX_SEG; ++s) { A.gseg[s] = gseg[s]; A.gmode[s] = gmode[s]; }
    auto al = [](int x, int q) { return (x + q - 1) / q * q; };
    const int NS = A.NS8 * 8, ND = A.ND8 * 8;
    if (NS > 256 || ND > 256) return SE3_OK;
    A.tmem_cols = 32;
    while (A.tmem_cols < NS + 4 * ND) A.tmem_cols <<= 1;
    if (A.tmem_cols > 512) return SE3_OK;
    const int nrg = TMI / 8;
    A.rg1 = ((A.N2 + A.N3) >> 2) * 128; A.rg2 = (A.N2 >> 2) * 128; A.rg3 = (A.N3 >> 2) * 128;
    A.sz_t1 = nrg * A.rg1; A.sz_t2 = nrg * A.rg2; A.sz_t3 = nrg * A.rg3;
    A.gts = (src.cum[src.nseg] + 3) & ~3;
    if ((A.gts & 31) == 0) A.gts += 4;
    int o = 0;
    A.o_t1 = o; o += 2 * A.sz_t1;
    A.o_t2 = o; o += 2 * A.sz_t2;
    A.o_t3 = o; o += 6 * A.sz_t3;
    A.o_b1 = o; o += 2 * NS * (A.N2 + A.N3) * 4;
    A.o_b2 = o; o += 2 * ND * A.N2 * 4;
    A.o_b3 = o; o += 2 * ND * A.N3 * 4;
    o = al(o, 128);
    // the input-gradient tile aliases the H tiles: it is written only after the MMAs of the tile have completed
    // (acc-full barrier) and the builders wait for the H-empty barrier (arrived after the scatter) before refilling
    if (TMI * A.gts * 4 <= 2 * A.sz_t1 + 2 * A.sz_t2 + 6 * A.sz_t3) A.o_gt = A.o_t1;
    else { A.o_gt = o; o += al(TMI * A.gts * 4, 16); }
    A.o_stage = o; o += 2 * A.slot_floats * 4;
    A.o_tab = o; o += al(ntab * 4, 16);
    A.o_norm = o; o += al((A.mz + 3 * A.mv) * 4, 16);
    A.o_tbl = o; o += al((8 * A.NSG8 + 8 * A.NDG8) * 4, 16);
    A.o_bar = o; o += 8 * 8 + 16;
    int dev = 0, maxsm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (o > maxsm) return SE3_OK;
    const int smem = std::max(o, 120 * 1024);
    static bool attr_set = false;
    if (!attr_set) {
        SE3_CUDA_TRY(cudaFuncSetAttribute(l1tp_tc_bwdi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
        attr_set = true;
    }
    const long long ntiles = (a->rows + TMI - 1) / TMI;
    const int grid = (int)std::min<long long>(ntiles, num_sms());
    l1tp_tc_bwdi_kernel<<<grid, TBI_THREADS, smem, st>>>(A);
    SE3_LAUNCHED();
    g_tc_launches.fetch_add(1, std::memory_order_relaxed);
    *launched = true;
    return SE3_OK;
}
