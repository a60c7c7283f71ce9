// Tensor-core (tcgen05 / TMEM) forward of the fused l<=1 tensor-product layer — the Blackwell-native
// version of the hot kernel for the SEGNN case (inputs a x0e + b x1o, outputs c x0e + d x1o).
//
// fp32 parity (1e-5) on the tensor pipe: every product is evaluated as 3xTF32,
//     a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi     (a_hi = fp32 truncated to tf32, a_lo = a - a_hi),
// three `tcgen05.mma.kind::tf32` per K-step with fp32 accumulation in TMEM.
//
// One persistent CTA per SM, 24 warps, warp-specialised, everything hand-shaken with mbarriers:
//   warps 0,2   producers : cp.async (16 B, zero-fill) gather of the tile's rows (virtual concat of indexed
//                           segments) + in2 into a 2-slot stage ring; index values prefetched one tile ahead
//   warps 4-15  builders  : stage -> feature chunks in the UMMA canonical K-major layout (64 rows x 24 K,
//                           hi and lo halves), chunk ring of <= 7 slots; fence.proxy.async; arrive
//   warp  1     MMA       : one lane issues 9 MMAs per chunk (3 K-steps x 3xTF32), tcgen05.commit frees the
//                           chunk slot; after the last chunk of a tile commits the accumulator buffer
//   warps 16-23 epilogue  : tcgen05.ld (32x32b) -> Y0/Y1 combination, norm, swish/sigmoid gate -> smem tile
//                           -> coalesced stores / residual / sorted-segment sum; TMEM accumulators are
//                           double buffered so the epilogue of tile t overlaps the MMAs of tile t+1
// GEMMs per 64-row tile (M=64; N and K padded with zero weights):
//   P[64 x N1] = S  [64 x K1] . [WZ_s | WV_s]      scalars, shared by the l=0 and l=1 outputs
//   Q[64 x N2] = Dd [64 x K2] . WZ_d               Dd = c3 <v, Y1>
//   Tc[64 x N3]= AVc[64 x K2] . WV_v  (c = x,y,z)  AVc = c3 Y0 v_c
//   out0[m] = norm (Y0 P[m] + Q[m]),   out1[m][c] = norm (c3 Y1[c] P[N2+m] + Tc[m])
#include <algorithm>

#include "tc_common.cuh"

namespace se3 {

static constexpr int TC_THREADS = 768;
static constexpr int TM = 64;
static constexpr int KC = 24;
static constexpr int SLOT_HALF = TM * KC * 4;  // bytes of one (hi or lo) 64x24 fp32 chunk
static constexpr int SLOT_BYTES = 2 * SLOT_HALF;
static constexpr int NBUILD_WARPS = 12;  // one 16-row x 2-K-chunk task per warp and chunk
static constexpr int NEPI_WARPS = 8;
static constexpr int BUILD_W0 = 4, EPI_W0 = 16;

struct TcArgs {
    long long rows;
    RowSrc src;
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
    const int* tab;
    int ns, nd, mz, mv, nchS, nchD, NCH, K1, K2, N1, N2, N3, nslot, acc_stride;
    int t_s, t_d, t_oz, t_ov, ntab;
    int d_out;
    int sstride[SE3_MAX_SEG], soff[SE3_MAX_SEG], vec16[SE3_MAX_SEG], swidth[SE3_MAX_SEG];
    int in2off, slot_floats;
    int dop, dpp;  // out / post tile strides
    // shared memory byte offsets
    int o_b1, o_b2, o_b3, o_a, o_stage, o_out, o_post, o_tab, o_norm, o_sq, o_bar;
};

__global__ void __launch_bounds__(TC_THREADS, 1) l1tp_tc_fwd_kernel(const TcArgs A) {
    extern __shared__ __align__(1024) unsigned char smraw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* smf = reinterpret_cast<float*>(smraw);
    int* tab = reinterpret_cast<int*>(smraw + A.o_tab);     // plan tables (virtual columns)
    float* norm = reinterpret_cast<float*>(smraw + A.o_norm);  // nz[mz], nv[3 mv]
    int* sq = reinterpret_cast<int*>(smraw + A.o_sq);       // per 16-byte scalar K-chunk: (seg<<16)|col or -1
    int* stab = sq + 6 * A.nchS;                            // per scalar k: (seg<<16)|col
    int* vtab = stab + A.K1;                                // per vector kd: (seg<<16)|col
    int* pcol = vtab + A.K2;                                // per post column: raw column of the value
    int* gcol = pcol + A.epi.d_post;                        // per post column: raw column of its gate (-1: swish)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smraw + A.o_bar);
    // barrier map
    const uint32_t bar0 = smem_u32(bars);
    auto BAR = [&](int i) { return bar0 + 8u * i; };
    const int B_STAGE_FULL = 0, B_STAGE_EMPTY = 2, B_ACC_FULL = 4, B_ACC_EMPTY = 6, B_SLOT_FULL = 8;
    const int B_SLOT_EMPTY = B_SLOT_FULL + A.nslot;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + (B_SLOT_EMPTY + A.nslot));

    // ---------------- one-time setup
    for (int t = tid; t < A.ntab; t += TC_THREADS) tab[t] = A.tab[t];
    for (int t = tid; t < A.mz; t += TC_THREADS) norm[t] = A.nz ? A.nz[t] : 1.0f;
    for (int t = tid; t < 3 * A.mv; t += TC_THREADS) norm[A.mz + t] = A.nv ? A.nv[t] : 1.0f;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(BAR(B_STAGE_FULL + i), 64);
            mbar_init(BAR(B_STAGE_EMPTY + i), NBUILD_WARPS);
            mbar_init(BAR(B_ACC_FULL + i), 1);
            mbar_init(BAR(B_ACC_EMPTY + i), NEPI_WARPS);
        }
        for (int i = 0; i < A.nslot; ++i) {
            mbar_init(BAR(B_SLOT_FULL + i), NBUILD_WARPS);
            mbar_init(BAR(B_SLOT_EMPTY + i), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // virtual column -> (segment, column) tables
    for (int k = tid; k < A.K1; k += TC_THREADS) {
        int e = -1;
        if (k < A.ns) {
            const int vc = tab[A.t_s + k];
            int s = 0;
            for (int q = 1; q < SE3_MAX_SEG; ++q)
                if (q < A.src.nseg && vc >= A.src.cum[q]) s = q;
            e = (A.sstride[s] << 20) | (A.soff[s] + vc - A.src.cum[s]);   // stride | float offset in the stage slot
        }
        stab[k] = e;
    }
    for (int k = tid; k < A.K2; k += TC_THREADS) {
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
    if (A.epi.mode == SE3_EPI_GATE) {
        for (int pp = tid; pp < A.epi.d_post; pp += TC_THREADS) {
            if (pp < A.epi.ns_g) {
                pcol[pp] = tab[A.t_oz + pp];
                gcol[pp] = -1;
            } else {
                const int qv = pp - A.epi.ns_g, v = qv / 3, c = qv - 3 * v;
                pcol[pp] = tab[A.t_ov + v] + c;
                gcol[pp] = tab[A.t_oz + A.epi.ns_g + v];
            }
        }
    }
    __syncthreads();
    for (int q = tid; q < 6 * A.nchS; q += TC_THREADS) {
        // 16-byte fast path: four consecutive, 16-byte aligned columns of one segment
        const int e0 = stab[4 * q];
        int f = -1;
        if (4 * q + 3 < A.ns && e0 >= 0 && (e0 & 3) == 0 && stab[4 * q + 1] == e0 + 1 && stab[4 * q + 2] == e0 + 2 &&
            stab[4 * q + 3] == e0 + 3)
            f = e0;
        sq[q] = f;
    }
    // weights -> canonical B tiles (hi | lo)
    {
        const int KQ1 = A.K1 >> 2, KQ2 = A.K2 >> 2;
        unsigned char* b1 = smraw + A.o_b1;
        const int half1 = A.N1 * A.K1 * 4;
        for (int t = tid; t < A.N1 * A.K1; t += TC_THREADS) {
            const int n = t / A.K1, k = t - n * A.K1;
            float x = 0.0f;
            if (k < A.ns) {
                if (n < A.mz) x = __ldg(A.wz + (long long)k * A.mz + n);
                else if (n >= A.N2 && n - A.N2 < A.mv) x = __ldg(A.wv + (long long)k * A.mv + (n - A.N2));
            }
            float hi, lo;
            split_tf32(x, hi, lo);
            const int o = canon_off(n, k, KQ1);
            *reinterpret_cast<float*>(b1 + o) = hi;
            *reinterpret_cast<float*>(b1 + half1 + o) = lo;
        }
        unsigned char* b2 = smraw + A.o_b2;
        const int half2 = A.N2 * A.K2 * 4;
        for (int t = tid; t < A.N2 * A.K2; t += TC_THREADS) {
            const int n = t / A.K2, k = t - n * A.K2;
            float x = 0.0f;
            if (k < A.nd && n < A.mz) x = __ldg(A.wz + (long long)(A.ns + k) * A.mz + n);
            float hi, lo;
            split_tf32(x, hi, lo);
            const int o = canon_off(n, k, KQ2);
            *reinterpret_cast<float*>(b2 + o) = hi;
            *reinterpret_cast<float*>(b2 + half2 + o) = lo;
        }
        unsigned char* b3 = smraw + A.o_b3;
        const int half3 = A.N3 * A.K2 * 4;
        for (int t = tid; t < A.N3 * A.K2; t += TC_THREADS) {
            const int n = t / A.K2, k = t - n * A.K2;
            float x = 0.0f;
            if (k < A.nd && n < A.mv) x = __ldg(A.wv + (long long)(A.ns + k) * A.mv + n);
            float hi, lo;
            split_tf32(x, hi, lo);
            const int o = canon_off(n, k, KQ2);
            *reinterpret_cast<float*>(b3 + o) = hi;
            *reinterpret_cast<float*>(b3 + half3 + o) = lo;
        }
    }
    fence_proxy_async();
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const long long R = A.rows;
    const long long ntiles = (R + TM - 1) / TM;

    if (warp == 0 || warp == 2) {
        // ================= producers: one row per lane
        const int prow = (warp == 0 ? 0 : 32) + lane;
        long long cur[SE3_MAX_SEG];
        {
            const long long gr = (long long)blockIdx.x * TM + prow;
#pragma unroll
            for (int s = 0; s < SE3_MAX_SEG; ++s)
                cur[s] = (s < A.src.nseg && gr < R) ? (A.src.idx[s] ? (long long)A.src.idx[s][gr] : gr) : 0;
        }
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int slot = it & 1, use = it >> 1;
            const long long gr = tile * TM + prow;
            const bool valid = gr < R;
            mbar_wait(BAR(B_STAGE_EMPTY + slot), (use & 1) ^ 1);
            const uint32_t sbase = smem_u32(smf) + A.o_stage + (uint32_t)slot * A.slot_floats * 4;
#pragma unroll
            for (int s = 0; s < SE3_MAX_SEG; ++s) {
                if (s >= A.src.nseg) break;
                const float* srcp = A.src.base[s] + (valid ? cur[s] * A.src.ld[s] : 0);
                const uint32_t dst = sbase + (A.soff[s] + prow * A.sstride[s]) * 4;
                const int w = A.swidth[s];
                if (A.vec16[s]) {
                    for (int c = 0; c < w; c += 4) cp_async16(dst + c * 4, srcp + c, valid);
                } else {
                    for (int c = 0; c < w; ++c) cp_async4(dst + c * 4, srcp + c, valid);
                }
            }
            cp_async16(sbase + (A.in2off + prow * 4) * 4, A.in2 + (valid ? gr * 4 : 0), valid);
            cp_async_mbar_arrive_noinc(BAR(B_STAGE_FULL + slot));
            // prefetch the index values of the next tile while this one is in flight
            const long long grn = (tile + gridDim.x) * TM + prow;
#pragma unroll
            for (int s = 0; s < SE3_MAX_SEG; ++s)
                cur[s] = (s < A.src.nseg && grn < R) ? (A.src.idx[s] ? (long long)A.src.idx[s][grn] : grn) : 0;
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else if (warp == 1) {
        // ================= MMA issuer
        const uint32_t id1 = make_idesc(A.N1), id2 = make_idesc(A.N2), id3 = make_idesc(A.N3);
        const uint32_t a_base = smem_u32(smraw) + A.o_a;
        const uint32_t b1 = smem_u32(smraw) + A.o_b1, b2 = smem_u32(smraw) + A.o_b2, b3 = smem_u32(smraw) + A.o_b3;
        const uint32_t half1 = A.N1 * A.K1 * 4, half2 = A.N2 * A.K2 * 4, half3 = A.N3 * A.K2 * 4;
        const uint32_t sbo1 = (A.K1 >> 2) * 128, sbo2 = (A.K2 >> 2) * 128;
        int it = 0;
        long long g = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int b = it & 1;
            mbar_wait(BAR(B_ACC_EMPTY + b), ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + (uint32_t)(b * A.acc_stride);
            for (int c = 0; c < A.NCH; ++c, ++g) {
                const int aslot = (int)(g % A.nslot);
                const uint32_t ause = (uint32_t)(g / A.nslot);
                mbar_wait(BAR(B_SLOT_FULL + aslot), ause & 1);
                tc_fence_after();
                if (lane == 0) {
                    uint32_t dcol, idesc, bb, bhalf, sbo;
                    int kofs, first;
                    if (c < A.nchS) {
                        dcol = 0; idesc = id1; bb = b1; bhalf = half1; sbo = sbo1; kofs = KC * c; first = (c == 0);
                    } else {
                        const int v = c - A.nchS, cd = v >> 2, sub = v & 3;
                        kofs = KC * cd; first = (cd == 0);
                        if (sub == 0) { dcol = A.N1; idesc = id2; bb = b2; bhalf = half2; sbo = sbo2; }
                        else { dcol = A.N1 + A.N2 + (sub - 1) * A.N3; idesc = id3; bb = b3; bhalf = half3; sbo = sbo2; }
                    }
                    const uint32_t ahi = a_base + (uint32_t)aslot * SLOT_BYTES, alo = ahi + SLOT_HALF;
#pragma unroll
                    for (int j = 0; j < KC / 8; ++j) {
                        const uint64_t da_hi = make_desc(ahi + j * 256, (KC / 4) * 128);
                        const uint64_t da_lo = make_desc(alo + j * 256, (KC / 4) * 128);
                        const uint32_t bo = (uint32_t)((kofs + 8 * j) >> 2) * 128;
                        const uint64_t db_hi = make_desc(bb + bo, sbo);
                        const uint64_t db_lo = make_desc(bb + bhalf + bo, sbo);
                        tc_mma_tf32(acc + dcol, da_hi, db_hi, idesc, (first && j == 0) ? 0u : 1u);
                        tc_mma_tf32(acc + dcol, da_hi, db_lo, idesc, 1u);
                        tc_mma_tf32(acc + dcol, da_lo, db_hi, idesc, 1u);
                    }
                    tc_commit(BAR(B_SLOT_EMPTY + aslot));
                }
                __syncwarp();
            }
            if (lane == 0) tc_commit(BAR(B_ACC_FULL + b));
            __syncwarp();
        }
    } else if (warp >= BUILD_W0 && warp < BUILD_W0 + NBUILD_WARPS) {
        // ================= builders
        const int bw = warp - BUILD_W0;
        const int row = (bw / 3) * 16 + (lane & 15), kc = (bw % 3) * 2 + (lane >> 4);
        const int o = (((row >> 3) * (KC / 4) + kc) << 7) + ((row & 7) << 4);  // byte offset inside a chunk half
        int it = 0;
        long long g0 = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it, g0 += A.NCH) {
            const int slot = it & 1, use = it >> 1;
            mbar_wait(BAR(B_STAGE_FULL + slot), use & 1);
            const float* st = smf + (A.o_stage >> 2) + (size_t)slot * A.slot_floats;
            // ---- scalar chunks
            for (int c = 0; c < A.nchS; ++c) {
                const long long g = g0 + c;
                const int aslot = (int)(g % A.nslot);
                mbar_wait(BAR(B_SLOT_EMPTY + aslot), ((uint32_t)(g / A.nslot) & 1) ^ 1);
                unsigned char* ahi = smraw + A.o_a + aslot * SLOT_BYTES;
                {
                    const int q = 6 * c + kc;
                    float4 v;
                    const int f = sq[q];
                    if (f >= 0) {
                        v = *reinterpret_cast<const float4*>(st + (f & 0xfffff) + row * (f >> 20));
                    } else {
                        float e[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int ee = stab[4 * q + j];
                            e[j] = ee >= 0 ? st[(ee & 0xfffff) + row * (ee >> 20)] : 0.0f;
                        }
                        v = make_float4(e[0], e[1], e[2], e[3]);
                    }
                    float4 h, l;
                    split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
                    *reinterpret_cast<float4*>(ahi + o) = h;
                    *reinterpret_cast<float4*>(ahi + SLOT_HALF + o) = l;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(B_SLOT_FULL + aslot));
            }
            // ---- vector chunk groups: D, Vx, Vy, Vz
            for (int cd = 0; cd < A.nchD; ++cd) {
                const long long g = g0 + A.nchS + 4 * cd;
                unsigned char* sl[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const long long gu = g + u;
                    const int aslot = (int)(gu % A.nslot);
                    mbar_wait(BAR(B_SLOT_EMPTY + aslot), ((uint32_t)(gu / A.nslot) & 1) ^ 1);
                    sl[u] = smraw + A.o_a + aslot * SLOT_BYTES;
                }
                {
                    const float4 y = *reinterpret_cast<const float4*>(st + A.in2off + row * 4);
                    const float sy0 = C3f * y.x;
                    float d[4], a0[4], a1[4], a2[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int kd = KC * cd + 4 * kc + j;
                        const int ee = kd < A.K2 ? vtab[kd] : -1;
                        float vx = 0.f, vy = 0.f, vz = 0.f;
                        if (ee >= 0) {
                            const float* p = st + (ee & 0xfffff) + row * (ee >> 20);
                            vx = p[0]; vy = p[1]; vz = p[2];
                        }
                        d[j] = C3f * (vx * y.y + vy * y.z + vz * y.w);
                        a0[j] = sy0 * vx; a1[j] = sy0 * vy; a2[j] = sy0 * vz;
                    }
                    float4 h, l;
#define SE3_PUT(arr, base)                                                                                   \
    split_tf32(arr[0], h.x, l.x); split_tf32(arr[1], h.y, l.y); split_tf32(arr[2], h.z, l.z);                \
    split_tf32(arr[3], h.w, l.w);                                                                            \
    *reinterpret_cast<float4*>(base + o) = h;                                                                \
    *reinterpret_cast<float4*>(base + SLOT_HALF + o) = l;
                    SE3_PUT(d, sl[0])
                    SE3_PUT(a0, sl[1])
                    SE3_PUT(a1, sl[2])
                    SE3_PUT(a2, sl[3])
#undef SE3_PUT
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) mbar_arrive(BAR(B_SLOT_FULL + (int)((g + u) % A.nslot)));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(B_STAGE_EMPTY + slot));
        }
    } else if (warp >= EPI_W0) {
        // ================= epilogue (8 warps).  Warp w reads TMEM lanes 32(w%4)..; rows 16(w%4)..+15 live in its
        // lanes 0..15.  Warps 8-11 convert the l=0 accumulators, warps 12-15 the l=1 ones, into the raw tile;
        // then all 256 threads finish: coalesced raw store (+residual), gate -> post store / segment sum.
        const int e = warp & 3;
        const int h = (warp - EPI_W0) >> 2;
        const int ew = warp - EPI_W0;      // 0..7
        const int et = tid - EPI_W0 * 32;  // 0..255
        const bool rowlane = lane < 16;
        const int row = 16 * e + (lane & 15);
        float* otile = reinterpret_cast<float*>(smraw + A.o_out);
        float* ptile = reinterpret_cast<float*>(smraw + A.o_post);
        const bool gate = A.epi.mode == SE3_EPI_GATE;
        const int dout = A.d_out, dpost = A.epi.d_post;
        const float* nz = norm;
        const float* nv = norm + A.mz;
        const int* oz = tab + A.t_oz;
        const int* ov = tab + A.t_ov;
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int b = it & 1;
            const long long row0 = tile * TM;
            const int nvalid = (int)min((long long)TM, R - row0);
            const long long gr = row0 + row;
            float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rowlane && gr < R) y = __ldg(reinterpret_cast<const float4*>(A.in2) + gr);
            mbar_wait(BAR(B_ACC_FULL + b), (it >> 1) & 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + (uint32_t)(b * A.acc_stride) + ((uint32_t)(32 * e) << 16);
            float* orow = otile + row * A.dop;
            if (h == 0) {
                for (int m0 = 0; m0 < A.mz; m0 += 8) {
                    float p[8], q[8];
                    tc_ld8(acc + m0, p);
                    tc_ld8(acc + A.N1 + m0, q);
                    tc_wait_ld();
                    if (rowlane) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int m = m0 + j;
                            if (m < A.mz) orow[oz[m]] = nz[m] * fmaf(y.x, p[j], q[j]);
                        }
                    }
                }
            } else {
                for (int m0 = 0; m0 < A.mv; m0 += 8) {
                    float pv[8], t0[8], t1[8], t2[8];
                    tc_ld8(acc + A.N2 + m0, pv);
                    tc_ld8(acc + A.N1 + A.N2 + m0, t0);
                    tc_ld8(acc + A.N1 + A.N2 + A.N3 + m0, t1);
                    tc_ld8(acc + A.N1 + A.N2 + 2 * A.N3 + m0, t2);
                    tc_wait_ld();
                    if (rowlane) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int m = m0 + j;
                            if (m < A.mv) {
                                const float sp = C3f * pv[j];
                                float* o = orow + ov[m];
                                o[0] = nv[3 * m] * fmaf(sp, y.y, t0[j]);
                                o[1] = nv[3 * m + 1] * fmaf(sp, y.z, t1[j]);
                                o[2] = nv[3 * m + 2] * fmaf(sp, y.w, t2[j]);
                            }
                        }
                    }
                }
            }
            // TMEM buffer can be refilled now
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(B_ACC_EMPTY + b));
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // ---- finish (256 threads): raw tile -> global (+residual); gate -> post rows (global and/or post tile)
            const bool to_ptile = gate && A.seg_idx;
            if (A.out_raw) {
                float* dst = A.out_raw + row0 * dout;
                const float* res = A.resid ? A.resid + row0 * dout : nullptr;
                const int total = nvalid * dout;
                if (A.dop == dout) {
                    // the tile is the exact image of the global rows: linear 16-byte copy
                    const int n4 = total >> 2;
                    for (int t = et; t < n4; t += 256) {
                        float4 v = reinterpret_cast<const float4*>(otile)[t];
                        if (res) {
                            const float4 q = __ldg(reinterpret_cast<const float4*>(res) + t);
                            v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
                        }
                        reinterpret_cast<float4*>(dst)[t] = v;
                    }
                    for (int t = (n4 << 2) + et; t < total; t += 256) dst[t] = otile[t] + (res ? __ldg(res + t) : 0.0f);
                } else {
                    for (int r = ew; r < nvalid; r += NEPI_WARPS) {
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
                if ((dpost & 3) == 0) {
                    const int q4 = dpost >> 2;           // float4 groups per row
                    const int total4 = nvalid * q4;
                    for (int t = et; t < total4; t += 256) {
                        const int r = t / q4, c0 = (t - r * q4) << 2;
                        const float* orow2 = otile + r * A.dop;
                        float v[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float x = orow2[pcol[c0 + j]];
                            const int gc = gcol[c0 + j];
                            const float gx = gc < 0 ? x : orow2[gc];
                            v[j] = (gc < 0 ? A.epi.cs : A.epi.cg) * sigm(gx) * x;
                        }
                        const float4 o4 = make_float4(v[0], v[1], v[2], v[3]);
                        if (dstp) reinterpret_cast<float4*>(dstp)[t] = o4;
                        if (to_ptile) *reinterpret_cast<float4*>(ptile + r * A.dpp + c0) = o4;
                    }
                } else {
                    for (int r = ew; r < nvalid; r += NEPI_WARPS) {
                        const float* orow2 = otile + r * A.dop;
                        for (int c = lane; c < dpost; c += 32) {
                            const float x = orow2[pcol[c]];
                            const int gc = gcol[c];
                            const float v = gc < 0 ? A.epi.cs * x * sigm(x) : A.epi.cg * sigm(orow2[gc]) * x;
                            if (dstp) dstp[r * dpost + c] = v;
                            if (to_ptile) ptile[r * A.dpp + c] = v;
                        }
                    }
                }
            }
            if (A.seg_idx) {
                if (to_ptile) asm volatile("bar.sync 1, 256;" ::: "memory");
                const float* stile = gate ? ptile : otile;
                const int sstr = gate ? A.dpp : A.dop, swid = gate ? dpost : dout;
                // sorted-segment sum: thread = (column, part of the rows); run-length combine, one red per run
                int parts = 256 / swid;
                if (parts < 1) parts = 1;
                const int rpp = (TM + parts - 1) / parts;
                for (int item = et; item < swid * parts; item += 256) {
                    const int c = item % swid, qd = item / swid;
                    const int rbeg = qd * rpp;
                    const int rend = min(rbeg + rpp, nvalid);
                    if (rbeg >= rend) continue;
                    int cur = __ldg(A.seg_idx + row0 + rbeg);
                    float accv = 0.0f;
                    for (int r = rbeg; r < rend; ++r) {
                        const int k = __ldg(A.seg_idx + row0 + r);
                        if (k != cur) {
                            atomicAdd(A.out_seg + (long long)cur * swid + c, accv);
                            cur = k;
                            accv = 0.0f;
                        }
                        accv += stile[r * sstr + c];
                    }
                    atomicAdd(A.out_seg + (long long)cur * swid + c, accv);
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
    }
    // ---------------- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

}  // namespace se3

using namespace se3;

// Returns SE3_OK and launches, or SE3_ERR_INVALID (with no error string change) when the configuration is not
// eligible for the tensor-core path; the caller then uses the generic kernel.
int se3_l1tp_tc_try_forward(const int n[4], const int m[4], const int t_in[4], const int t_out[4], int ntab,
                            const int* d_tab, const se3_l1tp_fwd_args* a, const RowSrc& src, const EpiL& epi,
                            cudaStream_t st, bool* launched) {
    *launched = false;
    static int disabled = -1;
    if (disabled < 0) {
        const char* e = getenv("SE3_DISABLE_TC");
        disabled = (e && e[0] == '1') ? 1 : 0;
    }
    if (disabled) return SE3_OK;
    // family E only, both output kinds, scalars and vectors present
    if (n[1] || n[2] || m[1] || m[2]) return SE3_OK;
    const int ns = n[0], nd = n[3], mz = m[0], mv = m[3];
    if (ns < 1 || nd < 1 || mz < 1 || mv < 1) return SE3_OK;
    TcArgs A;
    memset(&A, 0, sizeof(A));
    A.ns = ns; A.nd = nd; A.mz = mz; A.mv = mv;
    A.nchS = (ns + KC - 1) / KC; A.nchD = (nd + KC - 1) / KC;
    A.NCH = A.nchS + 4 * A.nchD;
    A.K1 = KC * A.nchS; A.K2 = KC * A.nchD;
    A.N2 = (mz + 7) & ~7; A.N3 = (mv + 7) & ~7; A.N1 = A.N2 + A.N3;
    if (A.N1 > 256) return SE3_OK;
    A.acc_stride = (A.N1 + A.N2 + 3 * A.N3 + 31) & ~31;
    if (2 * A.acc_stride > 512) return SE3_OK;
    if (epi.mode == SE3_EPI_GATE && ((epi.ns_g & 7) != 0 && false)) return SE3_OK;
    A.rows = a->rows; A.src = src; A.in2 = a->in2; A.wz = a->w[0]; A.wv = a->w[3]; A.nz = a->norm[0]; A.nv = a->norm[3];
    A.epi = epi; A.out_raw = a->out_raw; A.out_post = a->out_post; A.resid = a->resid; A.seg_idx = a->seg_idx;
    A.out_seg = a->out_seg; A.tab = d_tab; A.ntab = ntab;
    A.t_s = t_in[0]; A.t_d = t_in[3]; A.t_oz = t_out[0]; A.t_ov = t_out[3];
    A.d_out = mz + 3 * mv;
    if (((uintptr_t)a->in2 & 15) != 0) return SE3_OK;
    // stage layout
    int off = 0;
    for (int s = 0; s < src.nseg; ++s) {
        const int w = src.cum[s + 1] - src.cum[s];
        A.swidth[s] = w;
        A.sstride[s] = tc_stage_stride(w);
        A.soff[s] = off;
        A.vec16[s] = ((w & 3) == 0 && (src.ld[s] & 3) == 0 && ((uintptr_t)src.base[s] & 15) == 0) ? 1 : 0;
        if (!A.vec16[s] && w > 16) return SE3_OK;  // scalar cp.async only for narrow segments
        off += TM * A.sstride[s];
    }
    A.in2off = off;
    off += TM * 4;
    A.slot_floats = off;
    // raw tile stride: 16 row-lanes write one column at a time -> need stride*r (mod 32) distinct for r < 16,
    // i.e. stride odd or == 2 (mod 4).  If d_out itself qualifies the tile is the exact image of the global rows.
    A.dop = ((A.d_out & 1) || (A.d_out & 3) == 2) ? A.d_out : A.d_out + 2;
    A.dpp = (epi.d_post & 3) == 0 ? tc_stage_stride(epi.d_post) : (epi.d_post | 1);
    auto al = [](int x, int q) { return (x + q - 1) / q * q; };
    int o = 0;
    A.o_b1 = o; o += 2 * A.N1 * A.K1 * 4;
    A.o_b2 = o; o += 2 * A.N2 * A.K2 * 4;
    A.o_b3 = o; o += 2 * A.N3 * A.K2 * 4;
    o = al(o, 128);
    A.o_stage = o; o += 2 * A.slot_floats * 4;
    o = al(o, 16);
    A.o_out = o; o += al(TM * A.dop * 4, 16);
    A.o_post = o; o += (epi.mode == SE3_EPI_GATE && a->seg_idx) ? al(TM * A.dpp * 4, 16) : 0;
    A.o_tab = o; o += al(ntab * 4, 16);
    A.o_norm = o; o += al((mz + 3 * mv) * 4, 16);
    A.o_sq = o; o += al((6 * A.nchS + A.K1 + A.K2 + 2 * epi.d_post) * 4, 16);
    const int fixed = o + 1024;  // + barriers (filled in below)
    int dev = 0, maxsm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    int nslot = std::min(A.NCH, (maxsm - fixed - 128) / SLOT_BYTES);
    if (nslot < std::min(A.NCH, 5)) return SE3_OK;  // the vector group needs 4 slots + 1 for overlap
    if (nslot > 16) nslot = 16;
    A.nslot = nslot;
    o = al(o, 128);
    A.o_a = o; o += nslot * SLOT_BYTES;
    A.o_bar = o; o += (8 + 2 * nslot) * 8 + 16;
    if (o > maxsm) return SE3_OK;
    // > half of the SM's shared memory: exactly one CTA per SM, so the 512-column TMEM allocation never contends
    const int smem = std::max(o, 120 * 1024);
    static bool attr_set = false;
    if (!attr_set) {
        SE3_CUDA_TRY(cudaFuncSetAttribute(l1tp_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm));
        attr_set = true;
    }
    const long long ntiles = (a->rows + TM - 1) / TM;
    const int grid = (int)std::min<long long>(ntiles, num_sms());
    l1tp_tc_fwd_kernel<<<grid, TC_THREADS, smem, st>>>(A);
    SE3_LAUNCHED();
    g_tc_launches.fetch_add(1, std::memory_order_relaxed);
    *launched = true;
    return SE3_OK;
}
