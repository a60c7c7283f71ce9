// l <= 2 fully connected O(3) tensor product (se3_o3tp_*): first CUDA path for BASELINE configs[2] / SURVEY 8f-3.
// One persistent CTA per resident slot walks row tiles; per tile and output irrep: coupling features into shared memory,
// then the weight contraction as a register-blocked fp32 SIMT GEMM out of shared memory.  The tile
// programs live in o3tp_body.inl (shared with the CPU emulation under tests/emu), the planning in o3tp_tables.h.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "o3tp_tables.h"
#include "o3tp_lin.h"
#include "o3tp_tc.h"

namespace {

using se3::set_error;

typedef float4 o3f4;
#define O3_DEV __device__ __forceinline__
#define O3_THREADS { const int tid = threadIdx.x; const int NT = blockDim.x;
#define O3_END } __syncthreads();
#define O3_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define O3_GW_ADD(S, p, v)               \
    do {                                 \
        if ((S).gw_global)               \
            atomicAdd((p), (v));         \
        else                             \
            *(p) += (v);                 \
    } while (0)
#define O3_I2F(i) __int_as_float(i)
#define O3_ACC_DECL float (&acc)[o3::MAXIO_GW][16]
#define O3_ACC(acc, slot, tid) acc[slot]
#define O3_GLOBAL_ADD(p, v) atomicAdd((p), (v))
#define O3_GLOBAL_ADD4(p, a, b, c, d) se3::red_add_v4((p), (a), (b), (c), (d))
#define O3_CP4(dst, src)                                                                                      \
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), \
                 "l"(src)                                                                                    \
                 : "memory")
#define O3_CP_COMMIT() asm volatile("cp.async.commit_group;" ::: "memory")
#define O3_CP_WAIT() asm volatile("cp.async.wait_all;" ::: "memory")
#define O3_MULHI(a, b) __umulhi((a), (b))
#define O3_NT_DECL
#define O3_LD4(p) (*reinterpret_cast<const float4*>(p))
#define O3_UNROLL _Pragma("unroll")
#define O3_UNROLL2 _Pragma("unroll 2")
#include "o3tp_cg_gen.inl"
#include "o3tp_body.inl"

constexpr int O3_NT = 32 * o3::NWARP;

__device__ __forceinline__ const int32_t* load_table(const int32_t* __restrict__ tab_g, int32_t* sm) {
    const int words = tab_g[o3::H_WORDS];
    for (int i = threadIdx.x; i < words; i += blockDim.x) sm[i] = tab_g[i];
    __syncthreads();
    return sm;
}

__global__ void __launch_bounds__(O3_NT) o3tp_fwd_kernel(const int32_t* __restrict__ tab_g, const O3Rows in1,
                                                         const float* __restrict__ in2, const float* __restrict__ w,
                                                         float* __restrict__ out, long long rows, int TE) {
    extern __shared__ __align__(16) int32_t o3_sm[];
    const int32_t* tab = load_table(tab_g, o3_sm);
    float* fl = reinterpret_cast<float*>(o3_sm + tab[o3::H_WORDS]);
    float* Ws = fl;
    fl += tab[o3::H_NWP];
    O3Fwd S;
    S.tab = tab; S.Ws = Ws; S.TE = TE;
    S.xs0 = fl; fl += TE * (tab[o3::H_D1] | 1);
    S.xs1 = fl; fl += TE * (tab[o3::H_D1] | 1);
    S.ys0 = fl; fl += TE * (tab[o3::H_D2] | 1);
    S.ys1 = fl; fl += TE * (tab[o3::H_D2] | 1);
    S.os = fl;
    const long long ntiles = (rows + TE - 1) / TE;
    if ((long long)blockIdx.x < ntiles)
        o3_fwd_load(S, 0, in1, in2, (long long)blockIdx.x * TE, (int)min((long long)TE, rows - (long long)blockIdx.x * TE),
                    threadIdx.x, blockDim.x);
    for (int io = 0; io < tab[o3::H_NIO]; ++io) {
        const int32_t* IO = tab + tab[o3::H_IO] + io * o3::IO_W;
        const int mul = IO[o3::IO_MUL], K = IO[o3::IO_K], mulp = IO[o3::IO_MULP];
        for (int idx = threadIdx.x; idx < K * mulp; idx += blockDim.x) {
            const int kk = idx / mulp, c = idx - kk * mulp;
            Ws[IO[o3::IO_WSOFF] + idx] = c < mul ? w[IO[o3::IO_WOFF] + kk * mul + c] : 0.f;
        }
    }
    __syncthreads();
    int buf = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
        const long long row0 = tile * TE, next = tile + gridDim.x;
        const int nrow = (int)min((long long)TE, rows - row0);
        const int nrow_next = next < ntiles ? (int)min((long long)TE, rows - next * TE) : 0;
        o3_fwd_tile(S, buf, in1, in2, out, row0, nrow, next * TE, nrow_next);
    }
}

__global__ void __launch_bounds__(O3_NT) o3tp_bwd_kernel(const int32_t* __restrict__ tab_g, const O3Rows in1,
                                                         const float* __restrict__ in2, const float* __restrict__ w,
                                                         const float* __restrict__ gout, const O3GRows gin1,
                                                         float* __restrict__ gin2, float* __restrict__ gw, long long rows,
                                                         int gw_global, int dbuf) {
    extern __shared__ __align__(16) int32_t o3_sm[];
    const int32_t* tab = load_table(tab_g, o3_sm);
    float* fl = reinterpret_cast<float*>(o3_sm + tab[o3::H_WORDS]);
    float* WT = fl;
    fl += tab[o3::H_NWT];
    float* scr = fl;  // 16-byte aligned: read with float4
    fl += 16 * O3_SCR_LD;
    float* gWs = gw_global ? gw : fl;
    if (!gw_global) fl += tab[o3::H_NW];
    constexpr int TE = o3::TE_BWD;
    const int D1p = tab[o3::H_D1] | 1, D2p = tab[o3::H_D2] | 1, DOp = tab[o3::H_DOUT] | 1;
    O3Bwd S;
    S.tab = tab; S.WT = WT; S.gWs = gWs; S.gw_global = gw_global;
    S.xs0 = fl; fl += TE * D1p;
    S.xs1 = dbuf ? fl : S.xs0; fl += dbuf ? TE * D1p : 0;
    S.gxs = fl; fl += TE * D1p;
    S.ys0 = fl; fl += TE * D2p;
    S.ys1 = dbuf ? fl : S.ys0; fl += dbuf ? TE * D2p : 0;
    S.gys = fl; fl += TE * D2p;
    S.gs0 = fl; fl += TE * DOp;
    S.gs1 = dbuf ? fl : S.gs0; fl += dbuf ? TE * DOp : 0;
    S.F = fl; fl += (size_t)4 * tab[o3::H_MAXNP] * o3::NWARP * tab[o3::H_FROW];
    S.GT = fl;
    S.scr = scr;
    for (int io = 0; io < tab[o3::H_NIO]; ++io) {
        const int32_t* IO = tab + tab[o3::H_IO] + io * o3::IO_W;
        const int32_t* BL = tab + tab[o3::H_BLK] + IO[o3::IO_BLK];
        const int32_t* SUB = tab + tab[o3::H_SUB] + IO[o3::IO_SUB];
        const int mul = IO[o3::IO_MUL], KPP = 4 * IO[o3::IO_NSUB];
        for (int idx = threadIdx.x; idx < mul * KPP; idx += blockDim.x) {
            const int wi = idx / KPP, kkp = idx - wi * KPP, word = SUB[kkp >> 2];
            const int32_t* B = BL + (word & 0xffff) * o3::BLK_W;
            const int32_t* G = tab + tab[o3::H_GRP] + (B[o3::B_GRP] & 0xffff) * o3::GRP_W;
            const int32_t* P = tab + tab[o3::H_PATH] + G[o3::G_P0 + (word >> 16)] * o3::PATH_W;
            const int u = (B[o3::B_GRP] >> 16) + (kkp & 3);
            WT[IO[o3::IO_WTOFF] + idx] = u < G[o3::G_MUL1] ? w[P[o3::P_WOFF] + u * mul + wi] : 0.f;
        }
    }
    if (!gw_global)
        for (int idx = threadIdx.x; idx < tab[o3::H_NW]; idx += blockDim.x) gWs[idx] = 0.f;
    __syncthreads();
    const long long ntiles = (rows + TE - 1) / TE;
    if (dbuf && (long long)blockIdx.x < ntiles)
        o3_bwd_load(S, 0, in1, in2, gout, (long long)blockIdx.x * TE,
                    (int)min((long long)TE, rows - (long long)blockIdx.x * TE), threadIdx.x, blockDim.x);
    int buf = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= dbuf) {
        const long long row0 = tile * TE, next = tile + gridDim.x;
        const int nrow = (int)min((long long)TE, rows - row0);
        const int nrow_next = dbuf && next < ntiles ? (int)min((long long)TE, rows - next * TE) : 0;
        if (!dbuf) o3_bwd_load(S, 0, in1, in2, gout, row0, nrow, threadIdx.x, blockDim.x);  // single buffer: no overlap
        o3_bwd_tile(S, buf, in1, in2, gout, gin1, gin2, row0, nrow, next * TE, nrow_next);
    }
    __syncthreads();
    if (!gw_global)
        for (int idx = threadIdx.x; idx < tab[o3::H_NW]; idx += blockDim.x) atomicAdd(gw + idx, gWs[idx]);
}

__global__ void __launch_bounds__(O3_NT, 3) o3tp_gin_kernel(const int32_t* __restrict__ tab_g, const O3Rows in1,
                                                            const float* __restrict__ in2, const float* __restrict__ w,
                                                            const float* __restrict__ gout, const O3GRows gin1,
                                                            float* __restrict__ gin2, long long rows) {
    extern __shared__ __align__(16) int32_t o3_sm[];
    const int32_t* tab = load_table(tab_g, o3_sm);
    float* fl = reinterpret_cast<float*>(o3_sm + tab[o3::H_WORDS]);
    float* WT = fl;
    fl += tab[o3::H_NWT];
    constexpr int TE = o3::TE_GIN;
    const int D1p = tab[o3::H_D1] | 1, D2p = tab[o3::H_D2] | 1, DOp = tab[o3::H_DOUT] | 1;
    O3Gin S;
    S.tab = tab; S.WT = WT; S.need_gy = gin2 != nullptr;
    S.xs = fl; fl += TE * D1p;
    S.gxs = fl; fl += TE * D1p;
    S.ys = fl; fl += TE * D2p;
    S.gys = fl; fl += TE * D2p;
    S.gs = fl;
    for (int io = 0; io < tab[o3::H_NIO]; ++io) {   // a * W^T in sub-block order
        const int32_t* IO = tab + tab[o3::H_IO] + io * o3::IO_W;
        const int32_t* BL = tab + tab[o3::H_BLK] + IO[o3::IO_BLK];
        const int32_t* SUB = tab + tab[o3::H_SUB] + IO[o3::IO_SUB];
        const int mul = IO[o3::IO_MUL], KPP = 4 * IO[o3::IO_NSUB];
        const float a = __int_as_float(IO[o3::IO_A]);
        for (int idx = threadIdx.x; idx < mul * KPP; idx += blockDim.x) {
            const int wi = idx / KPP, kkp = idx - wi * KPP, word = SUB[kkp >> 2];
            const int32_t* B = BL + (word & 0xffff) * o3::BLK_W;
            const int32_t* G = tab + tab[o3::H_GRP] + (B[o3::B_GRP] & 0xffff) * o3::GRP_W;
            const int32_t* P = tab + tab[o3::H_PATH] + G[o3::G_P0 + (word >> 16)] * o3::PATH_W;
            const int u = (B[o3::B_GRP] >> 16) + (kkp & 3);
            WT[IO[o3::IO_WTOFF] + idx] = u < G[o3::G_MUL1] ? a * w[P[o3::P_WOFF] + u * mul + wi] : 0.f;
        }
    }
    __syncthreads();
    const long long ntiles = (rows + TE - 1) / TE;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long row0 = tile * TE;
        o3_gin_tile(S, in1, in2, gout, gin1, gin2, row0, (int)min((long long)TE, rows - row0));
    }
}

__global__ void __launch_bounds__(O3_NT, 2) o3tp_gw_kernel(const int32_t* __restrict__ tab_g, const O3Rows in1,
                                                           const float* __restrict__ in2, const float* __restrict__ gout,
                                                           float* __restrict__ gw, long long rows) {
    extern __shared__ __align__(16) int32_t o3_sm[];
    const int32_t* tab = load_table(tab_g, o3_sm);
    float* fl = reinterpret_cast<float*>(o3_sm + tab[o3::H_WORDS]);
    constexpr int TE = o3::TE_BWD;
    O3Gw S;
    S.tab = tab;
    S.F = fl; fl += max(tab[o3::H_FMAX], 16 * O3_SCR_LD);
    S.GT = fl; fl += tab[o3::H_GTMAX];
    S.xs = fl; fl += TE * (tab[o3::H_D1] | 1);
    S.ys = fl; fl += TE * (tab[o3::H_D2] | 1);
    S.gs = fl;
    float acc[o3::MAXIO_GW][16];
#pragma unroll
    for (int i = 0; i < o3::MAXIO_GW; ++i)
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[i][k] = 0.f;
    const long long ntiles = (rows + TE - 1) / TE;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long row0 = tile * TE;
        o3_gw_tile(S, acc, in1, in2, gout, row0, (int)min((long long)TE, rows - row0));
    }
    o3_gw_flush(S, acc, gw);
}

constexpr size_t SMEM_MAX = 227 * 1024;
constexpr size_t SMEM_TWO = 110 * 1024;  // budget that lets two CTAs share an SM

}  // namespace

struct se3_o3tp_plan {
    o3::Plan P;
    int32_t* d_tab = nullptr;
    int te_f = 0, te_b = 0, gw_global = 0, dbuf_b = 1, split = 0, grid_gin = 0, grid_gw = 0;
    size_t smem_gin = 0, smem_gw = 0;
    size_t smem_f = 0, smem_b = 0;
    int grid_f = 0, grid_b = 0;
    O3TcGw* tcgw = nullptr;   // weight gradient on the tensor cores (o3tp_tc_gw.cu), nullptr if the plan is not covered
    int gin_alone = 0;        // the input-gradient kernel can run next to it even when the SIMT pair is not used
    O3Lin* lin = nullptr;     // scalar second input: forward / input gradients as per-irrep linear maps (o3tp_lin.cu)
};

static int pick_tile(const std::vector<int32_t>& blob, bool bwd, int* te, size_t* smem) {
    const int cand[4] = {bwd ? o3::TE_BWD : 64, bwd ? 0 : 32, 0, 0};
    auto bytes = [&](int t) { return 4 * (blob.size() + (bwd ? o3::bwd_floats(blob) : o3::fwd_floats(blob, t))); };
    for (int c : cand)
        if (c && bytes(c) <= SMEM_TWO) { *te = c; *smem = bytes(c); return 0; }
    for (int c : cand)
        if (c && bytes(c) <= SMEM_MAX) { *te = c; *smem = bytes(c); return 0; }
    return SE3_ERR_TOO_LARGE;
}

extern "C" int se3_o3tp_plan_create(const se3_o3tp_desc* d, se3_o3tp_plan** out) {
    if (!d || !out) { set_error("null argument"); return SE3_ERR_INVALID; }
    *out = nullptr;
    if (d->n_in1 < 1 || d->n_in1 > SE3_O3_MAX_IRREPS || d->n_out < 1 || d->n_out > SE3_O3_MAX_IRREPS || d->n_in2 < 1 ||
        d->n_in2 > 3) {
        set_error("o3tp: irreps counts out of range (in1/out 1..%d, in2 1..3)", SE3_O3_MAX_IRREPS);
        return SE3_ERR_INVALID;
    }
    se3_o3tp_plan* p = new se3_o3tp_plan();
    for (int i = 0; i < d->n_in1; ++i) p->P.in1.push_back({d->in1_mul[i], d->in1_l[i], d->in1_p[i]});
    for (int i = 0; i < d->n_in2; ++i) p->P.in2.push_back({1, d->in2_l[i], d->in2_p[i]});
    for (int i = 0; i < d->n_out; ++i) p->P.out.push_back({d->out_mul[i], d->out_l[i], d->out_p[i]});
    if (!o3::build_plan(p->P)) {
        set_error("%s", p->P.err.c_str());
        delete p;
        return SE3_ERR_INVALID;
    }
    // the forward schedule depends on the tile size and adds a few table words: try 64 rows, fall back to 32
    o3::schedule_forward(p->P, 64);
    int rc_f = pick_tile(p->P.blob, false, &p->te_f, &p->smem_f);
    if (rc_f == 0 && p->te_f != 64) {
        o3::schedule_forward(p->P, p->te_f);
        rc_f = pick_tile(p->P.blob, false, &p->te_f, &p->smem_f);
        if (rc_f == 0 && p->te_f != p->P.blob[o3::H_TEF]) rc_f = SE3_ERR_TOO_LARGE;
    }
    // backward: two resident CTAs per SM matter more than the prefetch (measured: 6.0 vs 9.5 ms on the 64-wide update
    // product), so prefer in this order: double-buffered & two CTAs, single-buffered & two CTAs, double-buffered,
    // single-buffered, and last the weight-gradient accumulators in global memory (large irreps).
    int rc_b = SE3_ERR_TOO_LARGE;
    p->te_b = o3::TE_BWD;
    const struct { bool res, dbuf; size_t cap; } tries[5] = {
        {true, true, SMEM_TWO}, {true, false, SMEM_TWO}, {true, true, SMEM_MAX}, {true, false, SMEM_MAX}, {false, false, SMEM_MAX}};
    for (const auto& t : tries) {
        const size_t bytes = 4 * (p->P.blob.size() + o3::bwd_floats(p->P.blob, t.res, t.dbuf));
        if (bytes <= t.cap) {
            p->smem_b = bytes; p->gw_global = t.res ? 0 : 1; p->dbuf_b = t.dbuf ? 1 : 0;
            rc_b = 0;
            break;
        }
    }
    if (rc_f || rc_b) {
        set_error("o3tp: irreps too large for the shared-memory tiling (%d weights, d_in1 %d)", p->P.nW, p->P.D1);
        delete p;
        return SE3_ERR_TOO_LARGE;
    }
    cudaError_t e = cudaMalloc(&p->d_tab, p->P.blob.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_tab, p->P.blob.data(), p->P.blob.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(o3tp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(o3tp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
    // split backward (input gradients / weight gradients as two kernels) when the plan allows it and both fit twice per SM
    p->smem_gin = 4 * (p->P.blob.size() + o3::gin_floats(p->P.blob));
    p->smem_gw = 4 * (p->P.blob.size() + o3::gw_floats(p->P.blob));
    p->split = p->P.blob[o3::H_SPLIT] && p->smem_gin <= SMEM_TWO && p->smem_gw <= SMEM_TWO && !getenv("SE3_O3TP_FUSED_BWD");
    int bf = 0, bb = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bf, o3tp_fwd_kernel, O3_NT, p->smem_f);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bb, o3tp_bwd_kernel, O3_NT, p->smem_b);
    if (p->split) {
        int b1 = 0, b2 = 0;
        if (e == cudaSuccess) e = cudaFuncSetAttribute(o3tp_gin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(o3tp_gw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, o3tp_gin_kernel, O3_NT, p->smem_gin);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b2, o3tp_gw_kernel, O3_NT, p->smem_gw);
        if (b1 < 1 || b2 < 1) p->split = 0;
        p->grid_gin = b1 * se3::num_sms();
        p->grid_gw = b2 * se3::num_sms();
    }
    if (e != cudaSuccess || bf < 1 || bb < 1) {
        set_error("o3tp: CUDA setup failed: %s", e != cudaSuccess ? cudaGetErrorString(e) : "kernel does not fit an SM");
        if (p->d_tab) cudaFree(p->d_tab);
        delete p;
        return e != cudaSuccess ? (int)e : SE3_ERR_TOO_LARGE;
    }
    p->grid_f = bf * se3::num_sms();
    p->grid_b = bb * se3::num_sms();
    p->tcgw = o3tp_tc_gw_create(p->P);
    p->lin = o3lin_create(p->P);
    if (p->tcgw && !p->split && p->smem_gin <= SMEM_MAX) {
        int b1 = 0;
        cudaError_t e2 = cudaFuncSetAttribute(o3tp_gin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX);
        if (e2 == cudaSuccess) e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b1, o3tp_gin_kernel, O3_NT, p->smem_gin);
        if (e2 == cudaSuccess && b1 >= 1) { p->gin_alone = 1; p->grid_gin = b1 * se3::num_sms(); }
        else cudaGetLastError();
    }
    *out = p;
    return SE3_OK;
}

extern "C" void se3_o3tp_plan_destroy(se3_o3tp_plan* p) {
    if (!p) return;
    if (p->d_tab) cudaFree(p->d_tab);
    o3tp_tc_gw_destroy(p->tcgw);
    o3lin_destroy(p->lin);
    delete p;
}

extern "C" int se3_o3tp_plan_info(const se3_o3tp_plan* p, int32_t dims[8]) {
    if (!p || !dims) { set_error("null argument"); return SE3_ERR_INVALID; }
    dims[0] = p->P.D1; dims[1] = p->P.D2; dims[2] = p->P.Dout; dims[3] = (int32_t)p->P.paths.size();
    dims[4] = p->P.nW; dims[5] = p->te_f; dims[6] = p->te_b; dims[7] = (int32_t)(p->smem_b >> 10) | (p->gw_global << 16) | (p->dbuf_b << 17) | (p->split << 18) |
              ((p->tcgw && (p->split || p->gin_alone || p->lin) ? 1 : 0) << 19) | ((p->lin ? 1 : 0) << 20);
    return SE3_OK;
}

extern "C" int se3_o3tp_plan_paths(const se3_o3tp_plan* p, int32_t* i1, int32_t* i2, int32_t* io, int32_t* woff,
                                   float* pw) {
    if (!p) { set_error("null argument"); return SE3_ERR_INVALID; }
    for (size_t k = 0; k < p->P.paths.size(); ++k) {
        const o3::PathH& h = p->P.paths[k];
        if (i1) i1[k] = h.i1;
        if (i2) i2[k] = h.i2;
        if (io) io[k] = h.io;
        if (woff) woff[k] = h.woff;
        if (pw) pw[k] = p->P.a[h.io];
    }
    return SE3_OK;
}

extern "C" int se3_o3tp_coupling(int32_t l1, int32_t l2, int32_t l3, double* out) {
    double C[5][5][5];
    if (!out || !o3::cg(l1, l2, l3, C)) { set_error("o3tp: no coupling for (%d,%d,%d)", l1, l2, l3); return SE3_ERR_INVALID; }
    for (int i = 0; i < 2 * l1 + 1; ++i)
        for (int j = 0; j < 2 * l2 + 1; ++j)
            for (int k = 0; k < 2 * l3 + 1; ++k) *out++ = C[i][j][k];
    return SE3_OK;
}

static int make_rows(const se3_o3tp_plan* p, int nseg, const se3_rowseg* seg, O3Rows& X) {
    if (nseg < 1 || nseg > SE3_MAX_SEG || !seg) return SE3_ERR_INVALID;
    int c0 = 0;
    X.nseg = nseg;
    for (int s = 0; s < 4; ++s) {
        const bool on = s < nseg;
        if (on && (!seg[s].base || seg[s].width < 1 || seg[s].ld < seg[s].width)) return SE3_ERR_INVALID;
        X.base[s] = on ? seg[s].base : nullptr; X.idx[s] = on ? seg[s].idx : nullptr;
        X.ld[s] = on ? seg[s].ld : 0; X.c0[s] = c0; X.width[s] = on ? seg[s].width : 0;
        c0 += X.width[s];
    }
    return c0 == p->P.D1 ? SE3_OK : SE3_ERR_INVALID;
}

extern "C" int se3_o3tp_forward_seg(se3_o3tp_plan* p, int64_t rows, int32_t nseg, const se3_rowseg* seg, const float* in2,
                                    const float* w, float* out, void* stream) {
    O3Rows X;
    if (p && rows == 0) return SE3_OK;   // an empty slab of a decomposed run: empty tensors have NULL data pointers
    if (!p || rows < 0 || make_rows(p, nseg, seg, X) || (rows > 0 && (!in2 || !w || !out))) {
        set_error("o3tp forward: bad argument (segments must add up to d_in1)");
        return SE3_ERR_INVALID;
    }
    if (rows == 0) return SE3_OK;
    // scalar second input: the linear-map forward (o3tp_lin.cu) was measured slower than this kernel (2.9 vs 1.06 ms for
    // 1.1M rows of the node tables); its input-gradient kernel is used by the backward (3.2 vs 4.5 ms)
    if (p->lin && nseg == 1 && !X.idx[0] && X.ld[0] == p->P.D1 && getenv("SE3_O3TP_LIN_FWD"))
        return o3lin_forward(p->lin, rows, X.base[0], in2, w, out, (cudaStream_t)stream);
    const long long ntiles = (rows + p->te_f - 1) / p->te_f;
    const int grid = (int)std::min<long long>(ntiles, p->grid_f);
    o3tp_fwd_kernel<<<grid, O3_NT, p->smem_f, (cudaStream_t)stream>>>(p->d_tab, X, in2, w, out, rows, p->te_f);
    SE3_LAUNCHED();
    return SE3_OK;
}

extern "C" int se3_o3tp_backward_seg(se3_o3tp_plan* p, int64_t rows, int32_t nseg, const se3_rowseg* seg, const float* in2,
                                     const float* w, const float* gout, float* const* gseg, const int32_t* gseg_mode,
                                     float* gin2, float* gw, void* stream) {
    O3Rows X;
    if (p && rows == 0 && gw) {          // as above; the weight gradient of no rows is zero
        SE3_CUDA_TRY(cudaMemsetAsync(gw, 0, sizeof(float) * p->P.nW, (cudaStream_t)stream));
        return SE3_OK;
    }
    if (!p || rows < 0 || !gw || make_rows(p, nseg, seg, X) || !gseg || !gseg_mode || (rows > 0 && (!in2 || !w || !gout))) {
        set_error("o3tp backward: bad argument (segments must add up to d_in1)");
        return SE3_ERR_INVALID;
    }
    O3GRows G;
    G.nseg = nseg;
    for (int s = 0; s < 4; ++s) {
        const bool on = s < nseg && gseg[s] && gseg_mode[s] != SE3_GRAD_NONE;
        if (on && gseg_mode[s] != SE3_GRAD_STORE && gseg_mode[s] != SE3_GRAD_ATOMIC && gseg_mode[s] != SE3_GRAD_SORTED) {
            set_error("o3tp backward: unknown gradient mode");
            return SE3_ERR_INVALID;
        }
        if (on && gseg_mode[s] == SE3_GRAD_STORE && X.idx[s]) {
            set_error("o3tp backward: SE3_GRAD_STORE needs identity rows");
            return SE3_ERR_INVALID;
        }
        G.base[s] = on ? gseg[s] : nullptr; G.idx[s] = X.idx[s]; G.ld[s] = X.ld[s]; G.c0[s] = X.c0[s]; G.width[s] = X.width[s];
        G.mode[s] = !on ? 0 : (gseg_mode[s] == SE3_GRAD_STORE ? 1 : (gseg_mode[s] == SE3_GRAD_SORTED ? 3 : 2));
        if (on && G.mode[s] != 1 && (X.width[s] & 3) == 0 && (X.ld[s] & 3) == 0 && ((uintptr_t)gseg[s] & 15) == 0) G.mode[s] |= 16;
    }
    SE3_CUDA_TRY(cudaMemsetAsync(gw, 0, sizeof(float) * p->P.nW, (cudaStream_t)stream));
    if (rows == 0) return SE3_OK;
    // weight gradient on the tensor cores for the whole 32-row tiles of a dense in1 (o3tp_tc_gw.cu); the input gradients
    // stay on the SIMT kernel, the (< 32) remaining rows add their weight gradient through the SIMT path
    const long long rows_tc = rows & ~31ll;
    const bool dense = nseg == 1 && !X.idx[0] && X.ld[0] == p->P.D1;
    const bool lin_gin = p->lin && dense && !gin2 && ((G.mode[0] & 15) == 0 || (G.mode[0] & 15) == 1);
    if (p->tcgw && rows_tc > 0 && (p->split || p->gin_alone || lin_gin) && dense &&
        o3tp_tc_gw_aligned(p->tcgw, X.base[0], in2, gout)) {
        cudaStream_t st = (cudaStream_t)stream;
        if (lin_gin) {
            if ((G.mode[0] & 15) == 1) {
                const int rc = o3lin_gin(p->lin, rows, gout, in2, w, G.base[0], st);
                if (rc) return rc;
            }
        } else {
            const long long t1 = (rows + o3::TE_GIN - 1) / o3::TE_GIN;
            o3tp_gin_kernel<<<(int)std::min<long long>(t1, p->grid_gin), O3_NT, p->smem_gin, st>>>(p->d_tab, X, in2, w, gout, G, gin2, rows);
            SE3_LAUNCHED();
        }
        const int rc = o3tp_tc_gw_run(p->tcgw, rows_tc, X.base[0], in2, gout, gw, st);
        if (rc) return rc;
        const long long tail = rows - rows_tc;
        if (tail > 0) {
            O3Rows Xt = X;
            Xt.base[0] = X.base[0] + rows_tc * X.ld[0];
            const float* in2t = in2 + rows_tc * p->P.D2;
            const float* gt = gout + rows_tc * p->P.Dout;
            if (p->split) {
                o3tp_gw_kernel<<<1, O3_NT, p->smem_gw, st>>>(p->d_tab, Xt, in2t, gt, gw, tail);
            } else {
                O3GRows G0 = G;
                for (int s = 0; s < 4; ++s) G0.mode[s] = 0;
                o3tp_bwd_kernel<<<1, O3_NT, p->smem_b, st>>>(p->d_tab, Xt, in2t, w, gt, G0, nullptr, gw, tail, p->gw_global, p->dbuf_b);
            }
            SE3_LAUNCHED();
        }
        return SE3_OK;
    }
    if (p->split) {
        const long long t1 = (rows + o3::TE_GIN - 1) / o3::TE_GIN, t2 = (rows + o3::TE_BWD - 1) / o3::TE_BWD;
        o3tp_gin_kernel<<<(int)std::min<long long>(t1, p->grid_gin), O3_NT, p->smem_gin, (cudaStream_t)stream>>>(
            p->d_tab, X, in2, w, gout, G, gin2, rows);
        SE3_LAUNCHED();
        o3tp_gw_kernel<<<(int)std::min<long long>(t2, p->grid_gw), O3_NT, p->smem_gw, (cudaStream_t)stream>>>(
            p->d_tab, X, in2, gout, gw, rows);
        SE3_LAUNCHED();
        return SE3_OK;
    }
    const long long ntiles = (rows + p->te_b - 1) / p->te_b;
    const int grid = (int)std::min<long long>(ntiles, p->grid_b);
    o3tp_bwd_kernel<<<grid, O3_NT, p->smem_b, (cudaStream_t)stream>>>(p->d_tab, X, in2, w, gout, G, gin2, gw, rows,
                                                                     p->gw_global, p->dbuf_b);
    SE3_LAUNCHED();
    return SE3_OK;
}

extern "C" int se3_o3tp_forward(se3_o3tp_plan* p, int64_t rows, const float* in1, const float* in2, const float* w,
                                float* out, void* stream) {
    if (!p) { set_error("o3tp forward: bad argument"); return SE3_ERR_INVALID; }
    const se3_rowseg seg = {in1, nullptr, p->P.D1, p->P.D1};
    if (rows > 0 && !in1) { set_error("o3tp forward: bad argument"); return SE3_ERR_INVALID; }
    if (rows == 0) return rows < 0 ? SE3_ERR_INVALID : SE3_OK;
    return se3_o3tp_forward_seg(p, rows, 1, &seg, in2, w, out, stream);
}

extern "C" int se3_o3tp_backward(se3_o3tp_plan* p, int64_t rows, const float* in1, const float* in2, const float* w,
                                 const float* gout, float* gin1, float* gin2, float* gw, void* stream) {
    if (!p || !gw || (rows > 0 && (!in1 || !gin1))) { set_error("o3tp backward: bad argument"); return SE3_ERR_INVALID; }
    static const float dummy = 0.f;   // rows == 0: the segment still needs a non-null base
    const se3_rowseg seg = {in1 ? in1 : &dummy, nullptr, p->P.D1, p->P.D1};
    float* gs[SE3_MAX_SEG] = {gin1, nullptr, nullptr, nullptr};
    const int32_t mode[SE3_MAX_SEG] = {SE3_GRAD_STORE, 0, 0, 0};
    return se3_o3tp_backward_seg(p, rows, 1, &seg, in2, w, gout, gs, mode, gin2, gw, stream);
}
