// Probe of tcgen05.mma kind::tf32 operand conventions on sm_100a (no-swizzle canonical layouts):
//   (1) K-major A and B (known-good reference), (2) MN-major A and B taken from ROW-major tiles [k][mn],
//   (3) whether the tf32 operand conversion truncates or rounds the low 13 mantissa bits.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o mma_probe mma_probe.cu ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include "../../scalable-e3-gnn_b200/csrc/tc_common.cuh"
using namespace se3;

// tile: element (r, c) at ((r>>3)*CQ + (c>>2))*128 + (r&7)*16 + (c&3)*4   (rows r, columns c, CQ = cols/4)
__device__ __host__ inline int toff(int r, int c, int CQ) { return (((r >> 3) * CQ + (c >> 2)) << 7) + ((r & 7) << 4) + ((c & 3) << 2); }

// mode 0: D[m][n] = sum_k A[m][k] B[n][k]   A tile rows = m (64), cols = k (32); B tile rows = n (16), cols = k (32)   K-major
// mode 1: D[m][n] = sum_k A[k][m] B[k][n]   A tile rows = k (32), cols = m (M);  B tile rows = k (32), cols = n (16)   MN-major
__global__ void probe(const float* Ain, const float* Bin, float* Dout, int mode, int M, int N, int K, int lbo_variant) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* sa = sm;
    unsigned char* sbm = sm + 32768;
    const int ar = mode == 0 ? M : K, ac = mode == 0 ? K : M;
    const int br = mode == 0 ? N : K, bc = mode == 0 ? K : N;
    for (int t = tid; t < ar * ac; t += blockDim.x) { int r = t / ac, c = t % ac; *(float*)(sa + toff(r, c, ac / 4)) = Ain[t]; }
    for (int t = tid; t < br * bc; t += blockDim.x) { int r = t / bc, c = t % bc; *(float*)(sbm + toff(r, c, bc / 4)) = Bin[t]; }
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
        for (int ks = 0; ks < K / 8; ++ks) {
            uint64_t da, db;
            uint32_t id;
            if (mode == 0) {
                da = make_desc_ex(a0 + ks * 256, 128, (K / 4) * 128);
                db = make_desc_ex(b0 + ks * 256, 128, (K / 4) * 128);
                id = make_idesc_ex(M, N, 0, 0);
            } else {
                // MN-major: 16-byte groups of 4 MN elements 128 B apart (SBO), 8 k-rows 16 B apart inside a core
                // matrix, next 8 k-rows (next K-step) (cols/4)*128 B further
                const uint32_t lboA = lbo_variant ? 128 : (M / 4) * 128, lboB = lbo_variant ? 128 : (N / 4) * 128;
                const uint32_t sboA = lbo_variant ? (M / 4) * 128 : 128, sboB = lbo_variant ? (N / 4) * 128 : 128;
                da = make_desc_ex(a0 + ks * (M / 4) * 128, lboA, sboA);
                db = make_desc_ex(b0 + ks * (N / 4) * 128, lboB, sboB);
                id = make_idesc_ex(M, N, 1, 1);
            }
            tc_mma_tf32(tm, da, db, id, ks ? 1u : 0u);
        }
        tc_commit(smem_u32(&bar));
    }
    __syncthreads();
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    // dump all 128 lanes x N columns (4 warps)
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

int main() {
    const int K = 32, N = 16;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, 128 * 128 * 4); cudaMalloc(&dB, 128 * 128 * 4); cudaMalloc(&dD, 128 * 64 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int M : {64, 128}) {
        // logical operands: Aop[m][k], Bop[n][k]
        std::vector<float> Aop(M * K), Bop(N * K), ref(M * N, 0.f);
        for (int i = 0; i < M * K; ++i) Aop[i] = (float)((i * 7 + 3) % 13 - 6);
        for (int i = 0; i < N * K; ++i) Bop[i] = (float)((i * 5 + 1) % 11 - 5);
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += Aop[m * K + k] * Bop[n * K + k]; ref[m * N + n] = s; }
        for (int mode = 0; mode < 2; ++mode)
            for (int var = 0; var < (mode ? 2 : 1); ++var) {
                std::vector<float> Ain(M * K), Bin(N * K);
                if (mode == 0) { Ain = Aop; Bin = Bop; }
                else {
                    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) Ain[k * M + m] = Aop[m * K + k];
                    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) Bin[k * N + n] = Bop[n * K + k];
                }
                cudaMemcpy(dA, Ain.data(), M * K * 4, cudaMemcpyHostToDevice);
                cudaMemcpy(dB, Bin.data(), N * K * 4, cudaMemcpyHostToDevice);
                cudaMemset(dD, 0, 128 * 64 * 4);
                probe<<<1, 128, 65536>>>(dA, dB, dD, mode, M, N, K, var);
                cudaError_t e = cudaDeviceSynchronize();
                std::vector<float> D(128 * N);
                cudaMemcpy(D.data(), dD, 128 * N * 4, cudaMemcpyDeviceToHost);
                // row m of D lives in TMEM lane m (M=128) or lane 32*(m/16) + m%16 (M=64)
                int bad = 0;
                for (int m = 0; m < M; ++m) {
                    const int ln = M == 128 ? m : 32 * (m / 16) + (m % 16);
                    for (int n = 0; n < N; ++n) if (D[ln * N + n] != ref[m * N + n]) ++bad;
                }
                printf("M=%d mode=%s%s: %s (%d mismatches) err=%s  D[0][0..3]=%g %g %g %g ref=%g %g %g %g\n", M, mode ? "MN-major" : "K-major",
                       mode ? (var ? " (lbo/sbo swapped)" : "") : "", bad ? "FAIL" : "ok", bad, cudaGetErrorString(e), D[0], D[1], D[2], D[3],
                       ref[0], ref[1], ref[2], ref[3]);
            }
    }
    {   // rounding probe: A[0][0] = 1 + 2^-11 + 2^-12, B[0][0] = 1, everything else 0  (K-major, M=64)
        const int M = 64;
        std::vector<float> Aop(M * K, 0.f), Bop(N * K, 0.f);
        Aop[0] = 1.0f + ldexpf(1.0f, -11) + ldexpf(1.0f, -12);
        Bop[0] = 1.0f;
        Aop[K] = 1.0f + ldexpf(1.0f, -11);           // row 1: exactly half an ulp of tf32
        Bop[0] = 1.0f;
        cudaMemcpy(dA, Aop.data(), M * K * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, Bop.data(), N * K * 4, cudaMemcpyHostToDevice);
        probe<<<1, 128, 65536>>>(dA, dB, dD, 0, M, N, K, 0);
        cudaDeviceSynchronize();
        std::vector<float> D(128 * N);
        cudaMemcpy(D.data(), dD, 128 * N * 4, cudaMemcpyDeviceToHost);
        printf("tf32 conversion: 1+2^-11+2^-12 -> %.10f, 1+2^-11 -> %.10f  (truncate: 1.0 / 1.0; round-nearest: %.10f / tie)\n", D[0], D[N],
               1.0f + ldexpf(1.0f, -10));
    }
    return 0;
}
