// Shared PTX wrappers and layout helpers of the tcgen05 kernels (sm_100a).
#pragma once
#include "common.cuh"

namespace se3 {

static constexpr float C3f = 0.57735026918962576451f;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Waiting warps back off with nanosleep: a tight try_wait loop of many consumer warps was measured to take 43 % of all
// issued instructions of the weight-gradient kernel and to starve the two producer warps it was waiting for.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    while (!mbar_try(bar, parity)) __nanosleep(40);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// volatile loads: pinned in program order relative to the (volatile) mbarrier operations, so the compiler cannot sink a
// prefetch issued a tile ahead down to its first use
__device__ __forceinline__ float4 ldg4_v(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ldg2_v(const float* p) {
    float2 v;
    asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg1_v(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int ldgi_v(const int* p) {
    int v;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// 16 TMEM lanes x 8 columns per warp instruction, every lane gets data (mma.m16n8 accumulator fragment):
//   v[0], v[1] = (lane row g = laneid >> 2,     columns 2 (laneid & 3) + {0, 1})
//   v[2], v[3] = (lane row g + 8,               same columns)
// Used to drain M = 64 accumulators, whose rows occupy TMEM lanes 0..15 of every 32-lane quarter: the 32x32b shape
// leaves half of the warp without data.
__device__ __forceinline__ void tc_ld_16x256(uint32_t taddr, float (&v)[4]) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Sorted-segment sum of one 64-row tile (rows sorted by segment id): out[seg * ld + c] += sum of the tile rows of seg.
// Segments that lie strictly inside the tile have exactly one writer and are written with a plain store (the
// destination is zero-initialised); only the (at most two) segments shared with the neighbouring tiles use `red`.
//   pass A (all NT threads): thread = (column, 8-row part); complete runs inside a part -> store; the part's head and
//                            tail runs go to shared memory;
//   pass B (one thread per column): merge heads/tails across the parts in row order; store or red as above.
// sseg[0] = segment of the row before the tile (-1: none), sseg[1..64] = rows, sseg[65] = segment of the row after.
template <int NT>
__device__ __forceinline__ void sorted_segment_sum_tile(const float* tile, int tstride, int width, int nvalid, const int* sseg,
                                                        float* out, int ld, float4* hs, int tid, int barid) {
    int parts = NT / width;
    if (parts < 1) parts = 1;
    if (parts > 8) parts = 8;
    const int rpp = (64 + parts - 1) / parts;
    for (int item = tid; item < width * parts; item += NT) {
        const int c = item % width, qd = item / width;
        const int rbeg = qd * rpp, rend = min(rbeg + rpp, nvalid);
        float4 h = make_float4(__int_as_float(-1), 0.0f, __int_as_float(-1), 0.0f);
        if (rbeg < rend) {
            int cur = sseg[1 + rbeg];
            float acc = 0.0f, sum_h = 0.0f;
            const int seg_h = cur;
            bool first = true;
            for (int r = rbeg; r < rend; ++r) {
                const int k = sseg[1 + r];
                if (k != cur) {
                    if (first) { sum_h = acc; first = false; }
                    else out[(long long)cur * ld + c] = acc;
                    cur = k;
                    acc = 0.0f;
                }
                acc += tile[r * tstride + c];
            }
            if (first) h = make_float4(__int_as_float(seg_h), acc, __int_as_float(-1), 0.0f);
            else h = make_float4(__int_as_float(seg_h), sum_h, __int_as_float(cur), acc);
        }
        hs[qd * width + c] = h;
    }
    named_bar(barid, NT);
    // pass B, spread over all threads: a run is owned by the thread of the part it STARTS in (as that part's head, if the
    // previous non-empty part does not end with the same segment, or as its tail); the owner adds the heads of the
    // following parts that continue the run and emits it.
    const int prevseg = sseg[0], nextseg = sseg[65];
    for (int item = tid; item < width * parts; item += NT) {
        const int c = item % width, qd = item / width;
        const float4 h = hs[qd * width + c];
        const int sh = __float_as_int(h.x), st = __float_as_int(h.z);
        if (sh < 0) continue;
        auto emit = [&](int seg, float v) {
            float* p = out + (long long)seg * ld + c;
            if (seg == prevseg || seg == nextseg) atomicAdd(p, v);
            else *p = v;
        };
        auto extend = [&](int seg, float acc) {   // add the heads of the following parts while they continue `seg`
            for (int q2 = qd + 1; q2 < parts; ++q2) {
                const float4 g = hs[q2 * width + c];
                const int s2 = __float_as_int(g.x);
                if (s2 < 0) continue;
                if (s2 != seg) break;
                acc += g.y;
                if (__float_as_int(g.z) >= 0) break;   // that part goes on with another segment: the run ended there
            }
            emit(seg, acc);
        };
        bool head_owned = true;
        for (int q0 = qd - 1; q0 >= 0; --q0) {
            const float4 g = hs[q0 * width + c];
            const int s0 = __float_as_int(g.x);
            if (s0 < 0) continue;
            const int last = __float_as_int(g.z) >= 0 ? __float_as_int(g.z) : s0;
            head_owned = last != sh;
            break;
        }
        if (st >= 0) {
            if (head_owned) emit(sh, h.y);   // the head run ends inside this part
            extend(st, h.w);                 // the tail run starts here
        } else if (head_owned) {
            extend(sh, h.y);
        }
    }
}

// K-major, no-swizzle canonical layout: 8-row x 16-byte core matrices; LBO = 128 B between the two K-chunks
// of one MMA, SBO = bytes between consecutive 8-row groups (cute/arch/mma_sm100_desc.hpp, version 1).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((128u >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor: D=f32, A=B=tf32, both K-major, M=64
__device__ __forceinline__ uint32_t make_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
}
// general form: M in {64,128}, optional MN-major operands (bits 15 / 16)
__device__ __forceinline__ uint32_t make_idesc_ex(int m, int n, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// descriptor with explicit leading / stride byte offsets (no swizzle)
__device__ __forceinline__ uint64_t make_desc_ex(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}

// sigmoid with two MUFU ops (ex2.approx, rcp.approx: ~1e-7 relative, far inside the 1e-5 parity budget)
__device__ __forceinline__ float sigm(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}

// byte offset of element (row/col n, k) inside a canonical tile with KQ = K/4 sixteen-byte chunks per row
__device__ __forceinline__ int canon_off(int n, int k, int KQ) { return (((n >> 3) * KQ + (k >> 2)) << 7) + ((n & 7) << 4) + ((k & 3) << 2); }


// shared-memory row stride (floats) >= w with stride % 8 == 4: 8 consecutive rows read as float4 hit 8 distinct
// 4-bank groups, and rows stay 16-byte aligned
static inline int tc_stage_stride(int w) {
    int s = (w + 3) & ~3;
    while ((s & 7) != 4) s += 4;
    return s;
}

}  // namespace se3
