// Probe 2: tcgen05.mma kind::tf32 with MN-major operands in the SWIZZLE_128B_BASE32B layout (layout type 1):
// row-major [k][mn] tiles, 128-byte column chunks, 4-row atoms, 32-byte units XOR-swizzled with the row index.
// D[m][n] = sum_k A[k][m] B[k][n],  K = 32, M = 64, N = 32.
#include <cstdio>
#include <vector>
#include "../../scalable-e3-gnn_b200/csrc/tc_common.cuh"
using namespace se3;

__device__ __host__ inline int soff(int k, int mn, int K, int swz) {   // byte offset of element (k, mn)
    const int chunk = mn >> 5, c = mn & 31;
    int u = c >> 3;                       // 32-byte unit inside the 128-byte row
    if (swz) u ^= (k & 3);
    return chunk * K * 128 + k * 128 + u * 32 + (c & 7) * 4;
}
__device__ inline uint64_t mkdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t type) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           (1ull << 46) | ((uint64_t)type << 61);
}
__global__ void probe(const float* Ain, const float* Bin, float* Dout, int M, int N, int K, int swz, int swap, int type) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* sa = sm;
    unsigned char* sbm = sm + 32768;
    for (int t = tid; t < K * M; t += blockDim.x) { int k = t / M, m = t % M; *(float*)(sa + soff(k, m, K, swz)) = Ain[t]; }
    for (int t = tid; t < K * N; t += blockDim.x) { int k = t / N, n = t % N; *(float*)(sbm + soff(k, n, K, swz)) = Bin[t]; }
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    fence_proxy_async();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tslot;
    if (tid == 0) {
        const uint32_t a0 = smem_u32(sa), b0 = smem_u32(sbm);
        const uint32_t lbo = K * 128, sbo = 512;   // next 128-byte column chunk / next 4-row atom
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint64_t da = mkdesc(a0 + ks * 1024, swap ? sbo : lbo, swap ? lbo : sbo, type);
            const uint64_t db = mkdesc(b0 + ks * 1024, swap ? sbo : lbo, swap ? lbo : sbo, type);
            tc_mma_tf32(tm, da, db, make_idesc_ex(M, N, 1, 1), ks ? 1u : 0u);
        }
        tc_commit(smem_u32(&bar));
    }
    __syncthreads();
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    if (warp < 4) {
        for (int n0 = 0; n0 < N; n0 += 8) {
            float v[8];
            tc_ld8(tm + ((uint32_t)(32 * warp) << 16) + n0, v);
            tc_wait_ld();
            for (int j = 0; j < 8; ++j) Dout[(32 * warp + lane) * N + n0 + j] = v[j];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tm) : "memory");
}
__global__ void probe_off(const float* Ain, const float* Bin, float* Dout, int M, int K, int boff) {
    // B: 32-wide chunk, the N=16 operand occupies columns boff..boff+15; descriptor start address = chunk + boff*4 bytes
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* sa = sm;
    unsigned char* sbm = sm + 32768;
    for (int t = tid; t < K * 32 * 4; t += blockDim.x) ((float*)sbm)[t] = 1.0e30f;   // poison
    __syncthreads();
    for (int t = tid; t < K * M; t += blockDim.x) { int k = t / M, m = t % M; *(float*)(sa + soff(k, m, K, 1)) = Ain[t]; }
    for (int t = tid; t < K * 16; t += blockDim.x) { int k = t / 16, n = t % 16; *(float*)(sbm + soff(k, boff + n, K, 1)) = Bin[t]; }
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    fence_proxy_async();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tslot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tslot;
    if (tid == 0) {
        const uint32_t a0 = smem_u32(sa), b0 = smem_u32(sbm) + boff * 4;
        for (int ks = 0; ks < K / 8; ++ks)
            tc_mma_tf32(tm, mkdesc(a0 + ks * 1024, K * 128, 512, 1), mkdesc(b0 + ks * 1024, K * 128, 512, 1), make_idesc_ex(M, 16, 1, 1), ks ? 1u : 0u);
        tc_commit(smem_u32(&bar));
    }
    __syncthreads();
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    if (warp < 4) {
        for (int n0 = 0; n0 < 16; n0 += 8) {
            float v[8];
            tc_ld8(tm + ((uint32_t)(32 * warp) << 16) + n0, v);
            tc_wait_ld();
            for (int j = 0; j < 8; ++j) Dout[(32 * warp + lane) * 16 + n0 + j] = v[j];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tm) : "memory");
}
int main() {
    {
        const int K = 32, M = 64, N = 16;
        float *dA, *dB, *dD;
        cudaMalloc(&dA, 1 << 16); cudaMalloc(&dB, 1 << 16); cudaMalloc(&dD, 1 << 16);
        cudaFuncSetAttribute(probe_off, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        std::vector<float> A(K * M), B(K * N), ref(M * N, 0.f);
        for (int i = 0; i < K * M; ++i) A[i] = (float)((i * 7 + 3) % 13 - 6);
        for (int i = 0; i < K * N; ++i) B[i] = (float)((i * 5 + 1) % 11 - 5);
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += A[k * M + m] * B[k * N + n]; ref[m * N + n] = s; }
        cudaMemcpy(dA, A.data(), K * M * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), K * N * 4, cudaMemcpyHostToDevice);
        for (int boff : {0, 8, 16}) {
            probe_off<<<1, 128, 65536>>>(dA, dB, dD, M, K, boff);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> D(128 * N);
            cudaMemcpy(D.data(), dD, 128 * N * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int m = 0; m < M; ++m) { const int ln = 32 * (m / 16) + (m % 16); for (int n = 0; n < N; ++n) if (D[ln * N + n] != ref[m * N + n]) ++bad; }
            printf("B operand at column offset %d inside its chunk: %s (%d bad) %s\n", boff, bad ? "FAIL" : "OK", bad, cudaGetErrorString(e));
        }
    }

    const int K = 32;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, 1 << 16); cudaMalloc(&dB, 1 << 16); cudaMalloc(&dD, 1 << 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int M : {64, 128}) for (int N : {32, 64, 16}) {
        std::vector<float> A(K * M), B(K * N), ref(M * N, 0.f);
        for (int i = 0; i < K * M; ++i) A[i] = (float)((i * 7 + 3) % 13 - 6);
        for (int i = 0; i < K * N; ++i) B[i] = (float)((i * 5 + 1) % 11 - 5);
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += A[k * M + m] * B[k * N + n]; ref[m * N + n] = s; }
        cudaMemcpy(dA, A.data(), K * M * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), K * N * 4, cudaMemcpyHostToDevice);
        for (int type : {1, 2}) for (int swz = 0; swz < 2; ++swz) for (int swap = 0; swap < 2; ++swap) {
            cudaMemset(dD, 0, 1 << 16);
            probe<<<1, 128, 65536>>>(dA, dB, dD, M, N, K, swz, swap, type);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> D(128 * N);
            cudaMemcpy(D.data(), dD, 128 * N * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int m = 0; m < M; ++m) { const int ln = M == 128 ? m : 32 * (m / 16) + (m % 16); for (int n = 0; n < N; ++n) if (D[ln * N + n] != ref[m * N + n]) ++bad; }
            printf("M=%d N=%d type=%d swizzle=%d swapLS=%d: %s (%d bad) %s D0=%g %g %g ref=%g %g %g\n", M, N, type, swz, swap, bad ? "FAIL" : "OK", bad,
                   cudaGetErrorString(e), D[0], D[1], D[2], ref[0], ref[1], ref[2]);
        }
    }
    return 0;
}
