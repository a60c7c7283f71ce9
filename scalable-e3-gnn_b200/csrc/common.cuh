// Shared helpers for the se3gnn_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>

#include "../../include/se3gnn_b200.h"

namespace se3 {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;
extern std::atomic<long long> g_tc_launches;
int num_sms();

#define SE3_CUDA_TRY(expr)                                                        \
    do {                                                                          \
        cudaError_t _e = (expr);                                                  \
        if (_e != cudaSuccess) {                                                  \
            se3::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,          \
                           cudaGetErrorString(_e));                               \
            return (int)_e;                                                       \
        }                                                                         \
    } while (0)

#define SE3_LAUNCHED()                                                            \
    do {                                                                          \
        se3::g_launches.fetch_add(1, std::memory_order_relaxed);                  \
        SE3_CUDA_TRY(cudaGetLastError());                                         \
    } while (0)

__host__ __device__ __forceinline__ int pad4(int x) { return (x + 3) & ~3; }
// smallest 4*odd >= x (x>0); shared-memory row stride that keeps 16 B alignment and is
// conflict-free for 8 rows read as float4 by 8 lane groups.
__host__ __device__ __forceinline__ int stride4odd(int x) {
    int q = (x + 3) >> 2;
    if (q < 1) q = 1;
    if ((q & 1) == 0) q += 1;
    return q << 2;
}

__device__ __forceinline__ float f4c(const float4& v, int i) {
    return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

// red.global.add.v4.f32 (sm_90+): one 16-byte reduction instead of four.
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

// Row source of a fused TP call: in1 row = virtual concatenation of up to 4 (optionally indexed) segments.
struct RowSrc {
    const float* base[SE3_MAX_SEG];
    const int32_t* idx[SE3_MAX_SEG];
    int ld[SE3_MAX_SEG];
    int cum[SE3_MAX_SEG + 1];
    int nseg;
};

struct EpiL {
    int mode, ns_g, nv, d_post;
    float cs, cg;
};

}  // namespace se3
