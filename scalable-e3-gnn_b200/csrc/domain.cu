// Morton-range domain decomposition: what one rank keeps of the (replicated) global octree graph, computed on the GPU.
// Host logic and its CPU twin: se3gnn_b200/domain.py (builder-defined: the reference has no distributed code).
//
//   pass 1 (se3_domain_mark):  owner of every global node from the slab bounds (a cell belongs to the owner of its first
//          particle), exclusive scan of (owned, in-degree) -> local id of every owned node, local CSR row pointers,
//          counts {owned particles, owned nodes, local edges}                     -> ONE 24-byte D2H read sizes the arrays
//   pass 2 (se3_domain_edges): one warp per owned node: local (dst, src) ids of its CSR row in the global (dst, src)
//          order, its rows of the per-edge arrays (edge_attr 16 B, edge_extra 8 B), marks of the sources it does not own
//   pass 3 (se3_domain_halo):  scan of the marks -> halo nodes in ascending global id, stable grouping by owner rank
//          (the order the all-to-all-v needs), per-owner counts, and the halo sources of the local edges rewritten to
//          n_own + position                                                       -> ONE (8 + 8 world)-byte D2H read
// Everything is HBM-bound integer work: coalesced 4/8/16-byte accesses, no atomics except the per-owner counters.
#include <algorithm>

#include "common.cuh"

namespace se3 {

static constexpr int DSC = 1024;                       // elements per scan block

__device__ __forceinline__ int dd_owner(const long long* __restrict__ bounds, int world, long long first) {
    // last slab whose lower bound <= first (empty slabs never own anything); bounds is monotone, bounds[world] = n
    int lo = 0, hi = world - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(bounds + mid) <= first) lo = mid; else hi = mid - 1;
    }
    // skip empty slabs that share the lower bound: owner = LAST slab with bounds[r] <= first AND bounds[r+1] > first
    while (lo > 0 && __ldg(bounds + lo + 1) <= first) --lo;
    return lo;
}

__device__ __forceinline__ long long dd_first(long long g, long long n, const int* __restrict__ cell_start) {
    return g < n ? g : (long long)__ldg(cell_start + (g - n));
}

// packed value per element: owned << 40 | in-degree; per-block sums
__global__ void __launch_bounds__(256) dd_mark_kernel(long long n, long long nn, int rank, int world, const long long* __restrict__ bounds,
                                                       const int* __restrict__ cell_start, const long long* __restrict__ rowptr,
                                                       unsigned long long* __restrict__ packed, unsigned long long* __restrict__ bsum) {
    __shared__ unsigned long long wsum[8];
    const long long base = (long long)blockIdx.x * DSC;
    unsigned long long s = 0;
    for (int i = threadIdx.x; i < DSC; i += 256) {
        const long long g = base + i;
        unsigned long long v = 0;
        if (g < nn && dd_owner(bounds, world, dd_first(g, n, cell_start)) == rank)
            v = (1ull << 40) | (unsigned long long)(rowptr[g + 1] - rowptr[g]);
        if (g < nn) packed[g] = v;
        s += v;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; ++w) t += wsum[w];
        bsum[blockIdx.x] = t;
    }
}
// generic: per-block sums of a packed array that already exists (halo marks)
__global__ void __launch_bounds__(256) dd_bsum_kernel(const unsigned long long* __restrict__ packed, long long nn,
                                                       unsigned long long* __restrict__ bsum) {
    __shared__ unsigned long long wsum[8];
    const long long base = (long long)blockIdx.x * DSC;
    unsigned long long s = 0;
    for (int i = threadIdx.x; i < DSC; i += 256)
        if (base + i < nn) s += packed[base + i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; ++w) t += wsum[w];
        bsum[blockIdx.x] = t;
    }
}
// single block: exclusive scan of the block sums in place; total -> bsum[nb]
__global__ void __launch_bounds__(1024) dd_top_kernel(unsigned long long* __restrict__ bsum, long long nb) {
    __shared__ unsigned long long part[1024];
    const int t = threadIdx.x;
    const long long per = (nb + 1023) / 1024, a = min(nb, (long long)t * per), b = min(nb, a + per);
    unsigned long long s = 0;
    for (long long i = a; i < b; ++i) s += bsum[i];
    part[t] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned long long v = t >= o ? part[t - o] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    unsigned long long run = part[t] - s;
    for (long long i = a; i < b; ++i) { const unsigned long long v = bsum[i]; bsum[i] = run; run += v; }
    if (t == 1023) bsum[nb] = part[1023];
}
// exclusive scan inside every block (one thread per 4 consecutive elements, warp scans) + block offset
__device__ __forceinline__ unsigned long long dd_block_excl(unsigned long long mine, unsigned long long* wsum) {
    // exclusive prefix of `mine` over the 256 threads of the block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    unsigned long long off = 0;
    for (int w = 0; w < warp; ++w) off += wsum[w];
    __syncthreads();
    return off + inc - mine;
}
__global__ void __launch_bounds__(256) dd_own_final_kernel(const unsigned long long* __restrict__ packed, long long n, long long nn,
                                                            const unsigned long long* __restrict__ bsum, long long nb,
                                                            int* __restrict__ pos, long long* __restrict__ loc_rowptr,
                                                            long long* __restrict__ counts) {
    __shared__ unsigned long long wsum[8];
    const long long g0 = (long long)blockIdx.x * DSC + 4 * threadIdx.x;
    unsigned long long v[4], s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] = g0 + i < nn ? packed[g0 + i] : 0; s += v[i]; }
    unsigned long long run = bsum[blockIdx.x] + dd_block_excl(s, wsum);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long g = g0 + i;
        if (g <= nn) {
            const long long own_before = (long long)(run >> 40), e_before = (long long)(run & ((1ull << 40) - 1));
            if (g == n) counts[0] = own_before;                       // owned particles
            if (g == nn) { counts[1] = own_before; counts[2] = e_before; loc_rowptr[own_before] = e_before; }
            if (g < nn) {
                const bool own = (v[i] >> 40) != 0;
                pos[g] = own ? (int)own_before : -1;
                if (own) loc_rowptr[own_before] = e_before;
            }
        }
        run += v[i];
    }
}

// one warp per global node; owned nodes emit their CSR row in local numbering
__global__ void __launch_bounds__(256) dd_edges_kernel(long long nn, const int* __restrict__ pos, const long long* __restrict__ loc_rowptr,
                                                        const long long* __restrict__ rowptr, const int* __restrict__ col,
                                                        const float4* __restrict__ edge_attr, const float2* __restrict__ edge_extra,
                                                        int* __restrict__ own_ids, int* __restrict__ dst_loc, int* __restrict__ src_loc,
                                                        float4* __restrict__ attr_loc, float2* __restrict__ extra_loc,
                                                        unsigned long long* __restrict__ halo_mark) {
    const int lane = threadIdx.x & 31;
    const long long wstride = (long long)gridDim.x * 8;
    for (long long g = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); g < nn; g += wstride) {
        const int p = pos[g];
        if (p < 0) continue;
        if (lane == 0) own_ids[p] = (int)g;
        const long long e0 = rowptr[g], e1 = rowptr[g + 1], l0 = loc_rowptr[p];
        for (long long e = e0 + lane; e < e1; e += 32) {
            const long long k = l0 + (e - e0);
            const int s = col[e], ps = pos[s];
            dst_loc[k] = p;
            src_loc[k] = ps >= 0 ? ps : -1 - s;                 // halo sources are rewritten by dd_halo_fix_kernel
            if (ps < 0) halo_mark[s] = 1ull;                     // benign race: every writer stores the same value
            if (edge_attr) attr_loc[k] = edge_attr[e];
            if (edge_extra) extra_loc[k] = edge_extra[e];
        }
    }
}

// halo nodes in ascending global id: asc[rank among marks] = g; per-owner counts
__global__ void __launch_bounds__(256) dd_halo_final_kernel(const unsigned long long* __restrict__ mark, long long n, long long nn,
                                                             const unsigned long long* __restrict__ bsum, int world,
                                                             const long long* __restrict__ bounds, const int* __restrict__ cell_start,
                                                             int* __restrict__ asc, int* __restrict__ asc_owner,
                                                             long long* __restrict__ counts /* [3] = n_halo, [4..] per owner */) {
    __shared__ unsigned long long wsum[8];
    const long long g0 = (long long)blockIdx.x * DSC + 4 * threadIdx.x;
    unsigned long long v[4], s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] = g0 + i < nn ? mark[g0 + i] : 0; s += v[i]; }
    unsigned long long run = bsum[blockIdx.x] + dd_block_excl(s, wsum);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long g = g0 + i;
        if (g < nn && v[i]) {
            const int o = dd_owner(bounds, world, dd_first(g, n, cell_start));
            asc[run] = (int)g;
            asc_owner[run] = o;
            atomicAdd((unsigned long long*)(counts + 4 + o), 1ull);
        }
        if (g == nn - 1) counts[3] = (long long)(run + v[i]);
        run += v[i];
    }
}
// single block: stable grouping of the ascending halo list by owner; hpos[g] = position in the grouped list
__global__ void __launch_bounds__(1024) dd_halo_group_kernel(const int* __restrict__ asc, const int* __restrict__ asc_owner,
                                                              const long long* __restrict__ counts, int world,
                                                              int* __restrict__ halo_ids, int* __restrict__ hpos) {
    __shared__ int wsum[32];
    __shared__ int base_s;
    const int nh = (int)counts[3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int base = 0;
    for (int w = 0; w < world; ++w) {
        for (int c0 = 0; c0 < nh; c0 += 1024) {
            const int i = c0 + threadIdx.x;
            const bool f = i < nh && asc_owner[i] == w;
            const unsigned b = __ballot_sync(0xffffffffu, f);
            if (lane == 0) wsum[warp] = __popc(b);
            __syncthreads();
            int off = 0, tot = 0;
            for (int q = 0; q < 32; ++q) { const int c = wsum[q]; if (q < warp) off += c; tot += c; }
            if (f) {
                const int p = base + off + __popc(b & ((1u << lane) - 1));
                const int g = asc[i];
                halo_ids[p] = g;
                hpos[g] = p;
            }
            __syncthreads();
            base += tot;
        }
    }
    (void)base_s;
}
__global__ void dd_halo_fix_kernel(long long e_loc, int n_own, const int* __restrict__ hpos, int* __restrict__ src_loc) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= e_loc) return;
    const int v = src_loc[k];
    if (v < 0) src_loc[k] = n_own + hpos[-1 - v];
}

}  // namespace se3

using namespace se3;

extern "C" int se3_domain_work_bytes(int64_t nn, size_t* bytes) {
    if (nn < 0 || !bytes) { set_error("domain_work_bytes: bad argument"); return SE3_ERR_INVALID; }
    const size_t nb = (size_t)(nn + DSC) / DSC + 2;
    *bytes = ((size_t)nn + 1) * 8 /* packed */ + nb * 8 /* block sums */ + 256;
    return SE3_OK;
}

extern "C" int se3_domain_mark(int64_t n, int64_t m, int32_t rank, int32_t world, const int64_t* bounds,
                               const int32_t* cell_start, const int64_t* rowptr, int32_t* pos, int64_t* loc_rowptr,
                               int64_t* counts, void* work, size_t work_bytes, void* stream) {
    const long long nn = n + m;
    size_t need = 0;
    se3_domain_work_bytes(nn, &need);
    if (n < 1 || m < 0 || world < 1 || rank < 0 || rank >= world || !bounds || !rowptr || !pos || !loc_rowptr || !counts || !work ||
        work_bytes < need || (m > 0 && !cell_start)) {
        set_error("domain_mark: bad argument / workspace too small");
        return SE3_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* packed = (unsigned long long*)work;
    unsigned long long* bsum = packed + nn + 1;
    const long long nb = (nn + 1 + DSC - 1) / DSC;          // covers index nn (the totals)
    SE3_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(long long) * (4 + world), st));
    dd_mark_kernel<<<(unsigned)nb, 256, 0, st>>>(n, nn, rank, world, (const long long*)bounds, cell_start, (const long long*)rowptr, packed, bsum);
    SE3_LAUNCHED();
    dd_top_kernel<<<1, 1024, 0, st>>>(bsum, nb); SE3_LAUNCHED();
    dd_own_final_kernel<<<(unsigned)nb, 256, 0, st>>>(packed, n, nn, bsum, nb, pos, (long long*)loc_rowptr, (long long*)counts);
    SE3_LAUNCHED();
    return SE3_OK;
}

extern "C" int se3_domain_edges(int64_t n, int64_t m, int32_t world, const int64_t* bounds, const int32_t* cell_start,
                                const int32_t* pos, const int64_t* loc_rowptr, const int64_t* rowptr, const int32_t* col,
                                const float* edge_attr, const float* edge_extra, int64_t n_own, int64_t e_loc,
                                int32_t* own_ids, int32_t* dst_loc, int32_t* src_loc, float* attr_loc, float* extra_loc,
                                int32_t* halo_ids /*[nn] capacity*/, int32_t* hpos /*[nn]*/, int32_t* scratch /*[2 nn]*/,
                                int64_t* counts, void* work, size_t work_bytes, void* stream) {
    const long long nn = n + m;
    size_t need = 0;
    se3_domain_work_bytes(nn, &need);
    if (n < 1 || !bounds || !pos || !loc_rowptr || !rowptr || !counts || !work || work_bytes < need || !halo_ids || !hpos || !scratch ||
        n_own < 0 || e_loc < 0 || (e_loc > 0 && (!col || !dst_loc || !src_loc)) || (n_own > 0 && !own_ids)) {
        set_error("domain_edges: bad argument");
        return SE3_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* mark = (unsigned long long*)work;
    unsigned long long* bsum = mark + nn + 1;
    const long long nb = (nn + DSC - 1) / DSC;
    SE3_CUDA_TRY(cudaMemsetAsync(mark, 0, sizeof(unsigned long long) * (size_t)(nn + 1), st));
    const int grid = (int)std::max<long long>(1, std::min<long long>((nn + 7) / 8, (long long)num_sms() * 16));
    dd_edges_kernel<<<grid, 256, 0, st>>>(nn, pos, (const long long*)loc_rowptr, (const long long*)rowptr, col, (const float4*)edge_attr,
                                          (const float2*)edge_extra, own_ids, dst_loc, src_loc, (float4*)attr_loc, (float2*)extra_loc, mark);
    SE3_LAUNCHED();
    dd_bsum_kernel<<<(unsigned)nb, 256, 0, st>>>(mark, nn, bsum); SE3_LAUNCHED();
    dd_top_kernel<<<1, 1024, 0, st>>>(bsum, nb); SE3_LAUNCHED();
    int* asc = scratch;
    int* asc_owner = scratch + nn;
    dd_halo_final_kernel<<<(unsigned)nb, 256, 0, st>>>(mark, n, nn, bsum, world, (const long long*)bounds, cell_start, asc, asc_owner,
                                                       (long long*)counts);
    SE3_LAUNCHED();
    dd_halo_group_kernel<<<1, 1024, 0, st>>>(asc, asc_owner, (const long long*)counts, world, halo_ids, hpos); SE3_LAUNCHED();
    if (e_loc > 0) {
        dd_halo_fix_kernel<<<(unsigned)((e_loc + 255) / 256), 256, 0, st>>>(e_loc, (int)n_own, hpos, src_loc);
        SE3_LAUNCHED();
    }
    return SE3_OK;
}
