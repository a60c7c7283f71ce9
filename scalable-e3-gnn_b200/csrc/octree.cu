// Octree cell-splitting graph construction on the GPU (sm_100a).
//
// The reference's numba builder is not in the mount; the specification these
// kernels implement bit-for-bit is oracle/octree_oracle.py (self-authored from BASELINE.json's
// north_star).  Pipeline, all on one stream, two tiny D2H reads (cell count, edge count):
//   bbox (shuffle + ordered-int atomics) -> 63-bit Morton keys -> LSD radix sort, 8 x 8-bit passes,
//   stable (per-block histogram / row scan / warp-ballot multi-split scatter)
//   -> level-synchronous octree split: per level count -> ordered scan -> emit children (BFS ids)
//   -> leaf assignment -> 26-neighbour lookup by binary search in the level's sorted cell keys
//   -> degree scan -> coalesced CSR emission (one warp per cell) -> bottom-up cell moments
//   -> edge geometry (rel. position, SH(1), extras) and node attributes.
// Everything here is HBM-bound integer/byte work: no tensor cores, grids cover the data with
// coalesced 4/8/16-byte accesses.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace se3 {

static constexpr int MAXD = 21;
static constexpr int RS_THREADS = 256;
static constexpr int RS_ITEMS = 8;
static constexpr int RS_TILE = RS_THREADS * RS_ITEMS;

// ------------------------------------------------------------------ bbox + keys

__device__ __forceinline__ unsigned enc_f(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void bbox_init_kernel(unsigned* b) {
    if (threadIdx.x < 3) b[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6) b[threadIdx.x] = 0u;
}

__global__ void bbox_kernel(const float* __restrict__ pos, long long n, unsigned* b) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = __ldg(pos + 3 * i + a);
            mn[a] = fminf(mn[a], v);
            mx[a] = fmaxf(mx[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(b + a, enc_f(mn[a]));
            atomicMax(b + 3 + a, enc_f(mx[a]));
        }
    }
}

__device__ __forceinline__ unsigned long long spread3(unsigned long long v) {
    v &= 0x1FFFFFull;
    v = (v | (v << 32)) & 0x1F00000000FFFFull;
    v = (v | (v << 16)) & 0x1F0000FF0000FFull;
    v = (v | (v << 8)) & 0x100F00F00F00F00Full;
    v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}
__device__ __forceinline__ unsigned compact3(unsigned long long v) {
    v &= 0x1249249249249249ull;
    v = (v | (v >> 2)) & 0x10C30C30C30C30C3ull;
    v = (v | (v >> 4)) & 0x100F00F00F00F00Full;
    v = (v | (v >> 8)) & 0x1F0000FF0000FFull;
    v = (v | (v >> 16)) & 0x1F00000000FFFFull;
    v = (v | (v >> 32)) & 0x1FFFFFull;
    return (unsigned)v;
}

// bbox words: [0..2] lo (encoded), [3..5] hi (encoded); writes decoded lo + scale to fb[0..3]
__global__ void morton_kernel(const float* __restrict__ pos, long long n, const unsigned* __restrict__ b,
                              unsigned long long* __restrict__ keys, int* __restrict__ vals, float* fb) {
    const float lx = dec_f(b[0]), ly = dec_f(b[1]), lz = dec_f(b[2]);
    float L = fmaxf(fmaxf(__fsub_rn(dec_f(b[3]), lx), __fsub_rn(dec_f(b[4]), ly)), __fsub_rn(dec_f(b[5]), lz));
    if (!(L > 0.0f)) L = 1.0f;
    const float scale = __fdiv_rn(2097152.0f, L);
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { fb[0] = lx; fb[1] = ly; fb[2] = lz; fb[3] = scale; }
    if (i >= n) return;
    const float tx = __fmul_rn(__fsub_rn(__ldg(pos + 3 * i), lx), scale);
    const float ty = __fmul_rn(__fsub_rn(__ldg(pos + 3 * i + 1), ly), scale);
    const float tz = __fmul_rn(__fsub_rn(__ldg(pos + 3 * i + 2), lz), scale);
    const long long qmax = (1ll << MAXD) - 1;
    const unsigned long long qx = (unsigned long long)min((long long)floorf(tx), qmax);
    const unsigned long long qy = (unsigned long long)min((long long)floorf(ty), qmax);
    const unsigned long long qz = (unsigned long long)min((long long)floorf(tz), qmax);
    keys[i] = (spread3(qx) << 2) | (spread3(qy) << 1) | spread3(qz);
    vals[i] = (int)i;
}

// ------------------------------------------------------------------ radix sort (one 8-bit pass = 3 kernels)

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const unsigned long long* __restrict__ keys, long long n,
                                                              int shift, int nb, int* __restrict__ counts) {
    __shared__ int hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * RS_TILE;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const long long idx = base + i * RS_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&hist[(int)((keys[idx] >> shift) & 255ull)], 1);
    }
    __syncthreads();
    counts[(long long)threadIdx.x * nb + blockIdx.x] = hist[threadIdx.x];
}

__device__ __forceinline__ int block_excl_scan_256(int v, int* smem /*[8]*/, int& total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) smem[w] = x;
    __syncthreads();
    int wbase = 0, tot = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int s = smem[q];
        if (q < w) wbase += s;
        tot += s;
    }
    __syncthreads();
    total = tot;
    return wbase + x - v;
}

// one block per digit row: exclusive scan of counts[d][0..nb) in place, row total -> row_total[d]
__global__ void __launch_bounds__(256) rs_scan_rows_kernel(int* __restrict__ counts, int nb, int* __restrict__ row_total) {
    __shared__ int sm[8];
    int* row = counts + (long long)blockIdx.x * nb;
    int carry = 0;
    for (int c0 = 0; c0 < nb; c0 += 256) {
        const int i = c0 + threadIdx.x;
        const int v = i < nb ? row[i] : 0;
        int tot;
        const int e = block_excl_scan_256(v, sm, tot);
        if (i < nb) row[i] = carry + e;
        carry += tot;
    }
    if (threadIdx.x == 0) row_total[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const unsigned long long* __restrict__ kin,
                                                                 const int* __restrict__ vin,
                                                                 unsigned long long* __restrict__ kout,
                                                                 int* __restrict__ vout, long long n, int shift, int nb,
                                                                 const int* __restrict__ counts,
                                                                 const int* __restrict__ row_total) {
    __shared__ int whist[RS_THREADS / 32][256];
    __shared__ int goff[256];
    __shared__ int sm[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < RS_THREADS / 32; ++q) whist[q][threadIdx.x] = 0;
    int tot;
    const int dbase = block_excl_scan_256(row_total[threadIdx.x], sm, tot);
    goff[threadIdx.x] = dbase + counts[(long long)threadIdx.x * nb + blockIdx.x];
    __syncthreads();
    const long long base = (long long)blockIdx.x * RS_TILE + (long long)w * (32 * RS_ITEMS);
    unsigned long long key[RS_ITEMS];
    int val[RS_ITEMS], rank[RS_ITEMS];
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const long long idx = base + i * 32 + lane;
        const bool valid = idx < n;
        key[i] = valid ? kin[idx] : 0ull;
        val[i] = valid ? vin[idx] : 0;
        const int d = (int)((key[i] >> shift) & 255ull);
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        const unsigned peers = __match_any_sync(0xffffffffu, d) & vmask;
        int old = 0;
        if (valid) old = whist[w][d];
        __syncwarp();
        if (valid && lane == (__ffs(peers) - 1)) whist[w][d] = old + __popc(peers);
        __syncwarp();
        rank[i] = old + __popc(peers & lt);
    }
    __syncthreads();
    {
        int run = 0;
#pragma unroll
        for (int q = 0; q < RS_THREADS / 32; ++q) {
            const int t = whist[q][threadIdx.x];
            whist[q][threadIdx.x] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const long long idx = base + i * 32 + lane;
        if (idx < n) {
            const int d = (int)((key[i] >> shift) & 255ull);
            const long long p = (long long)goff[d] + whist[w][d] + rank[i];
            kout[p] = key[i];
            vout[p] = val[i];
        }
    }
}

// ------------------------------------------------------------------ level-synchronous split

struct TreeP {
    const unsigned long long* keys;
    long long n;
    int leaf_size, max_depth;
    long long cap;
    int *start, *count, *level, *parent, *first_child, *nchild;
    unsigned long long* ckey;
    int* level_ptr;  // [max_depth+2]
    int* bsum;       // block sums scratch
    int* flags;      // [0] overflow
};

__global__ void tree_init_kernel(TreeP T) {
    if (threadIdx.x == 0) {
        T.start[0] = 0; T.count[0] = (int)T.n; T.level[0] = 0; T.parent[0] = -1; T.first_child[0] = -1; T.nchild[0] = 0;
        T.ckey[0] = 0ull;
        T.level_ptr[0] = 0;
        for (int l = 1; l <= T.max_depth + 1; ++l) T.level_ptr[l] = 1;
        T.flags[0] = 0;
    }
}

__device__ __forceinline__ int lower_digit(const unsigned long long* keys, int lo, int hi, int shift, int o) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((int)((keys[mid] >> shift) & 7ull) < o) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ void split_count_block(const TreeP T, int lev, int vb, int* sm) {
    const int beg = T.level_ptr[lev], end = T.level_ptr[lev + 1];
    const int c = beg + vb * 256 + threadIdx.x;
    if (beg + vb * 256 >= end) return;
    int nc = 0;
    if (c < end && lev < T.max_depth) {
        const int cnt = T.count[c];
        if (cnt > T.leaf_size) {
            const int s = T.start[c], e = s + cnt, shift = 3 * (T.max_depth - lev - 1);
            int prev = s;
            for (int o = 1; o <= 8; ++o) {
                const int b = o < 8 ? lower_digit(T.keys, prev, e, shift, o) : e;
                nc += (b > prev);
                prev = b;
            }
        }
    }
    if (c < end) T.nchild[c] = nc;
    int tot;
    block_excl_scan_256(nc, sm, tot);
    if (threadIdx.x == 0) T.bsum[vb] = tot;
}
__global__ void __launch_bounds__(256) split_count_kernel(TreeP T, int lev) {
    __shared__ int sm[8];
    split_count_block(T, lev, blockIdx.x, sm);
}

__device__ __forceinline__ void split_scan_block(const TreeP T, int lev, int* sm) {
    const int beg = T.level_ptr[lev], end = T.level_ptr[lev + 1];
    const int nblk = (end - beg + 255) / 256;
    int carry = 0;
    for (int c0 = 0; c0 < nblk; c0 += 256) {
        const int i = c0 + threadIdx.x;
        const int v = i < nblk ? T.bsum[i] : 0;
        int tot;
        const int e = block_excl_scan_256(v, sm, tot);
        if (i < nblk) T.bsum[i] = carry + e;
        carry += tot;
    }
    if (threadIdx.x == 0) {
        if ((long long)end + carry > T.cap) {
            T.flags[0] = 1;  // capacity overflow: stop splitting
            carry = 0;
            T.flags[1] = lev;
        }
        for (int l = lev + 2; l <= T.max_depth + 1; ++l) T.level_ptr[l] = end + carry;
    }
}
__global__ void __launch_bounds__(256) split_scan_kernel(TreeP T, int lev) {
    __shared__ int sm[8];
    split_scan_block(T, lev, sm);
}

__device__ __forceinline__ void split_emit_block(const TreeP T, int lev, int vb, int* sm) {
    const int beg = T.level_ptr[lev], end = T.level_ptr[lev + 1];
    if (beg + vb * 256 >= end) return;
    if (T.level_ptr[lev + 2] == end) return;  // nothing split (or overflow)
    const int c = beg + vb * 256 + threadIdx.x;
    const int nc = c < end ? T.nchild[c] : 0;
    int tot;
    const int ex = block_excl_scan_256(nc, sm, tot);
    if (c >= end || nc == 0) return;
    int child = end + T.bsum[vb] + ex;
    T.first_child[c] = child;
    const int s = T.start[c], e = s + T.count[c], shift = 3 * (T.max_depth - lev - 1);
    int prev = s;
    for (int o = 1; o <= 8; ++o) {
        const int b = o < 8 ? lower_digit(T.keys, prev, e, shift, o) : e;
        if (b > prev) {
            T.start[child] = prev; T.count[child] = b - prev; T.level[child] = lev + 1; T.parent[child] = c;
            T.first_child[child] = -1; T.nchild[child] = 0;
            T.ckey[child] = T.keys[prev] >> shift;
            ++child;
        }
        prev = b;
    }
}

__global__ void __launch_bounds__(256) split_emit_kernel(TreeP T, int lev) {
    __shared__ int sm[8];
    split_emit_block(T, lev, blockIdx.x, sm);
}

// All levels in ONE cooperative launch: count / scan / emit per level separated by grid-wide barriers instead of three
// launches per level (63 launches of a few microseconds each dominated the build of a 100k-particle cloud), and the loop
// ends at the first level that holds no cell or splits nothing (every later level is empty by construction).  Blocks
// stride over the level's 256-cell groups, so the per-cell code is the one of the per-level kernels.
__global__ void __launch_bounds__(256) split_levels_kernel(TreeP T) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ int sm[8];
    for (int lev = 0; lev < T.max_depth; ++lev) {
        const int beg = T.level_ptr[lev], end = T.level_ptr[lev + 1];
        if (end <= beg) break;                          // grid-uniform: written before the last barrier
        const int nvb = (end - beg + 255) / 256;
        for (int vb = blockIdx.x; vb < nvb; vb += gridDim.x) {
            split_count_block(T, lev, vb, sm);
            __syncthreads();
        }
        grid.sync();
        if (blockIdx.x == 0) split_scan_block(T, lev, sm);
        grid.sync();
        if (T.level_ptr[lev + 2] == end) break;         // nothing split (or capacity overflow): grid-uniform
        for (int vb = blockIdx.x; vb < nvb; vb += gridDim.x) {
            split_emit_block(T, lev, vb, sm);
            __syncthreads();
        }
        grid.sync();
    }
}

__global__ void leaf_assign_kernel(TreeP T, const int* __restrict__ order, int* __restrict__ leaf_of_rank,
                                   int* __restrict__ cell_of_particle, int m) {
    // one warp per cell: lanes stride over the leaf's ranks
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= m || T.first_child[c] >= 0) return;
    const int s = T.start[c], cnt = T.count[c];
    for (int j = lane; j < cnt; j += 32) {
        leaf_of_rank[s + j] = c;
        cell_of_particle[order[s + j]] = c;
    }
}

// ------------------------------------------------------------------ neighbours, degrees, CSR emission

__global__ void cell_degree_kernel(TreeP T, int m, long long n, int* __restrict__ nbr, int* __restrict__ deg) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    const int lev = T.level[c];
    int found[26];
    int nf = 0;
    if (lev > 0) {
        const int beg = T.level_ptr[lev], end = T.level_ptr[lev + 1];
        const unsigned long long k = T.ckey[c];
        const int cx = (int)compact3(k >> 2), cy = (int)compact3(k >> 1), cz = (int)compact3(k);
        const int lim = 1 << lev;
        for (int dx = -1; dx <= 1; ++dx)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dz = -1; dz <= 1; ++dz) {
                    if (!(dx | dy | dz)) continue;
                    const int x = cx + dx, y = cy + dy, z = cz + dz;
                    if (x < 0 || y < 0 || z < 0 || x >= lim || y >= lim || z >= lim) continue;
                    const unsigned long long nk = (spread3((unsigned)x) << 2) | (spread3((unsigned)y) << 1) | spread3((unsigned)z);
                    int lo = beg, hi = end;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (T.ckey[mid] < nk) lo = mid + 1;
                        else hi = mid;
                    }
                    if (lo < end && T.ckey[lo] == nk) {
                        // insert ascending (ids are in key order within a level)
                        int p = nf++;
                        while (p > 0 && found[p - 1] > lo) { found[p] = found[p - 1]; --p; }
                        found[p] = lo;
                    }
                }
    }
    for (int i = 0; i < 26; ++i) nbr[(long long)c * 26 + i] = i < nf ? found[i] : -1;
    const bool leaf = T.first_child[c] < 0;
    deg[n + c] = (leaf ? T.count[c] : 0) + (T.parent[c] >= 0 ? 1 : 0) + nf + T.nchild[c];
}

__global__ void particle_degree_kernel(TreeP T, long long n, const int* __restrict__ leaf_of_rank, int* __restrict__ deg) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) deg[r] = T.count[leaf_of_rank[r]];
}

// generic exclusive scan int32 -> int64 (3 kernels), 1024 elements per block
__global__ void __launch_bounds__(256) scan_bsum_kernel(const int* __restrict__ in, long long n, long long* __restrict__ bsum) {
    __shared__ long long sm[8];
    const long long base = (long long)blockIdx.x * 1024 + threadIdx.x * 4;
    long long s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (base + i < n) s += in[base + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int q = 0; q < 8; ++q) t += sm[q];
        bsum[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024) scan_top_kernel(long long* __restrict__ bsum, long long nblk, long long* __restrict__ total) {
    __shared__ long long sm[32];
    __shared__ long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (long long c0 = 0; c0 < nblk; c0 += 1024) {
        const long long i = c0 + threadIdx.x;
        const long long v = i < nblk ? bsum[i] : 0;
        long long x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) sm[w] = x;
        __syncthreads();
        long long wbase = 0, tot = 0;
        for (int q = 0; q < 32; ++q) {
            const long long s = sm[q];
            if (q < w) wbase += s;
            tot += s;
        }
        const long long carry = carry_s;
        if (i < nblk) bsum[i] = carry + wbase + x - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

__global__ void __launch_bounds__(256) scan_final_kernel(const int* __restrict__ in, long long n,
                                                         const long long* __restrict__ bsum, long long* __restrict__ out) {
    __shared__ long long sm[8];
    const long long base = (long long)blockIdx.x * 1024 + threadIdx.x * 4;
    long long v[4], s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[i] = base + i < n ? in[base + i] : 0;
        s += v[i];
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long x = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) sm[w] = x;
    __syncthreads();
    long long wbase = 0;
    for (int q = 0; q < w; ++q) wbase += sm[q];
    long long run = bsum[blockIdx.x] + wbase + x - s;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
}

__global__ void __launch_bounds__(256) emit_edges_kernel(TreeP T, int m, long long n, const int* __restrict__ nbr,
                                                         const long long* __restrict__ rowptr, int* __restrict__ col,
                                                         int* __restrict__ dst) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= m) return;
    const bool leaf = T.first_child[c] < 0;
    const int s = T.start[c], cnt = T.count[c];
    const int cnode = (int)n + c;
    if (leaf) {
        // particle rows of this leaf are contiguous: cnt rows of cnt entries each
        const long long base = rowptr[s];
        const long long total = (long long)cnt * cnt;
        for (long long t = lane; t < total; t += 32) {
            const int row = (int)(t / cnt), j = (int)(t - (long long)row * cnt);
            col[base + t] = j == cnt - 1 ? cnode : s + j + (j >= row ? 1 : 0);
            dst[base + t] = s + row;
        }
    }
    const long long base = rowptr[n + c];
    const int deg = (int)(rowptr[n + c + 1] - base);
    const int np = leaf ? cnt : 0;
    const int par = T.parent[c];
    const int hasp = par >= 0 ? 1 : 0;
    const int nch = T.nchild[c];
    const int nn = deg - np - hasp - nch;
    const int fc = T.first_child[c];
    for (int t = lane; t < deg; t += 32) {
        int v;
        if (t < np) v = s + t;
        else {
            int u = t - np;
            if (hasp && u == 0) v = (int)n + par;
            else {
                u -= hasp;
                v = u < nn ? (int)n + nbr[(long long)c * 26 + u] : (int)n + fc + (u - nn);
            }
        }
        col[base + t] = v;
        dst[base + t] = cnode;
    }
}

// ------------------------------------------------------------------ node data

__global__ void permute_particles_kernel(long long n, const int* __restrict__ order, const float* __restrict__ pos,
                                         const float* __restrict__ vel, const float* __restrict__ mass,
                                         float* __restrict__ npos, float* __restrict__ nvel, float* __restrict__ nmass) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const long long i = order[r];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        npos[3 * r + a] = __ldg(pos + 3 * i + a);
        nvel[3 * r + a] = __ldg(vel + 3 * i + a);
    }
    nmass[r] = __ldg(mass + i);
}

__global__ void cell_moments_kernel(TreeP T, int lev, long long n, float* __restrict__ npos, float* __restrict__ nvel,
                                    float* __restrict__ nmass) {
    const int beg = T.level_ptr[lev], end = T.level_ptr[lev + 1];
    const int c = beg + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= end) return;
    float m = 0.f, px = 0.f, py = 0.f, pz = 0.f, vx = 0.f, vy = 0.f, vz = 0.f;
    int a, b;
    if (T.first_child[c] < 0) { a = T.start[c]; b = a + T.count[c]; }
    else { a = (int)n + T.first_child[c]; b = a + T.nchild[c]; }
    for (int j = a; j < b; ++j) {
        const float w = nmass[j];
        m += w;
        px += w * npos[3ll * j]; py += w * npos[3ll * j + 1]; pz += w * npos[3ll * j + 2];
        vx += w * nvel[3ll * j]; vy += w * nvel[3ll * j + 1]; vz += w * nvel[3ll * j + 2];
    }
    const float inv = m > 0.f ? 1.0f / m : 0.f;
    const long long o = n + c;
    nmass[o] = m;
    npos[3 * o] = px * inv; npos[3 * o + 1] = py * inv; npos[3 * o + 2] = pz * inv;
    nvel[3 * o] = vx * inv; nvel[3 * o + 1] = vy * inv; nvel[3 * o + 2] = vz * inv;
}

static constexpr float SH0 = 0.28209479177387814f;  // 1/(2 sqrt(pi))        ('integral' normalisation)
static constexpr float SH1 = 0.4886025119029199f;   // sqrt(3/(4 pi))

// per edge: rel = pos[src]-pos[dst]; edge_attr = (SH0, SH1*rel/|rel|); extra = (|rel|, mass_scale^2 m_i m_j)
__global__ void edge_geom_kernel(long long e, const int* __restrict__ dst, const int* __restrict__ col,
                                 const float* __restrict__ npos, const float* __restrict__ nmass, float mass_scale,
                                 float* __restrict__ eattr, float* __restrict__ extra) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e) return;
    const long long d = dst[i], s = col[i];
    const float rx = npos[3 * s] - npos[3 * d], ry = npos[3 * s + 1] - npos[3 * d + 1], rz = npos[3 * s + 2] - npos[3 * d + 2];
    const float r2 = rx * rx + ry * ry + rz * rz;
    const float r = sqrtf(r2);
    const float inv = r > 0.f ? SH1 / r : 0.f;
    reinterpret_cast<float4*>(eattr)[i] = make_float4(SH0, rx * inv, ry * inv, rz * inv);
    reinterpret_cast<float2*>(extra)[i] = make_float2(r, (mass_scale * nmass[d]) * (mass_scale * nmass[s]));
}

// per node: node_attr = mean of incoming edge_attr + SH(vel);  x_in = [pos-centroid, vel, |vel|, mass*mass_scale]
__global__ void node_feat_kernel(long long nn, long long n, const long long* __restrict__ rowptr,
                                 const float* __restrict__ eattr, const float* __restrict__ npos,
                                 const float* __restrict__ nvel, const float* __restrict__ nmass, float mass_scale,
                                 float* __restrict__ nattr, float* __restrict__ xin) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    const long long a = rowptr[i], b = rowptr[i + 1];
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (long long k = a; k < b; ++k) {
        const float4 v = reinterpret_cast<const float4*>(eattr)[k];
        s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w;
    }
    const float invd = b > a ? 1.0f / (float)(b - a) : 0.f;
    const float vx = nvel[3 * i], vy = nvel[3 * i + 1], vz = nvel[3 * i + 2];
    const float vn = sqrtf(vx * vx + vy * vy + vz * vz);
    const float iv = vn > 0.f ? SH1 / vn : 0.f;
    reinterpret_cast<float4*>(nattr)[i] = make_float4(s0 * invd + SH0, s1 * invd + vx * iv, s2 * invd + vy * iv, s3 * invd + vz * iv);
    // centroid = centre of mass of the root cell (node n)
    const float cx = npos[3 * n], cy = npos[3 * n + 1], cz = npos[3 * n + 2];
    float* x = xin + 8 * i;
    reinterpret_cast<float4*>(x)[0] = make_float4(npos[3 * i] - cx, npos[3 * i + 1] - cy, npos[3 * i + 2] - cz, vx);
    reinterpret_cast<float4*>(x)[1] = make_float4(vy, vz, vn, nmass[i] * mass_scale);
}

// ---- SH(2) attributes for the l_max = 2 tensor product (csrc/o3tp.cu): 9 columns = Y0 | Y1 (x,y,z) | Y2 in the
// orthonormal basis (xy, yz, 2zz-xx-yy, zx, xx-yy), 'integral' normalisation like SH0 / SH1 above
static constexpr float SH2 = 0.6307831305050401f * 1.2247448713915890f;  // sqrt(5/(4 pi)) * sqrt(3/2)

__device__ __forceinline__ void sh2_of(float x, float y, float z, float* o) {
    const float r = sqrtf(x * x + y * y + z * z);
    const float inv = r > 0.f ? 1.0f / r : 0.f, inv1 = r > 0.f ? SH1 / r : 0.f;   // inv1: same arithmetic as edge_geom_kernel
    const float nx = x * inv, ny = y * inv, nz = z * inv;
    o[0] = SH0;
    o[1] = x * inv1; o[2] = y * inv1; o[3] = z * inv1;
    o[4] = SH2 * 1.4142135623730951f * nx * ny;
    o[5] = SH2 * 1.4142135623730951f * ny * nz;
    o[6] = SH2 * 0.4082482904638631f * (2.f * nz * nz - nx * nx - ny * ny);
    o[7] = SH2 * 1.4142135623730951f * nz * nx;
    o[8] = SH2 * 0.7071067811865476f * (nx * nx - ny * ny);
}

__global__ void edge_sh2_kernel(long long e, const int* __restrict__ dst, const int* __restrict__ col,
                                const float* __restrict__ npos, float* __restrict__ eattr) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= e) return;
    const long long d = dst[i], s = col[i];
    float o[9];
    sh2_of(npos[3 * s] - npos[3 * d], npos[3 * s + 1] - npos[3 * d + 1], npos[3 * s + 2] - npos[3 * d + 2], o);
#pragma unroll
    for (int k = 0; k < 9; ++k) eattr[9 * i + k] = o[k];
}

// node_attr = mean of the incoming edge attributes + SH(2)(velocity)
__global__ void node_sh2_kernel(long long nn, const long long* __restrict__ rowptr, const float* __restrict__ eattr,
                                const float* __restrict__ nvel, float* __restrict__ nattr) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nn) return;
    const long long a = rowptr[i], b = rowptr[i + 1];
    float s[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) s[k] = 0.f;
    for (long long j = a; j < b; ++j)
#pragma unroll
        for (int k = 0; k < 9; ++k) s[k] += eattr[9 * j + k];
    const float invd = b > a ? 1.0f / (float)(b - a) : 0.f;
    float o[9];
    sh2_of(nvel[3 * i], nvel[3 * i + 1], nvel[3 * i + 2], o);
#pragma unroll
    for (int k = 0; k < 9; ++k) nattr[9 * i + k] = s[k] * invd + o[k];
}

}  // namespace se3

using namespace se3;

static inline unsigned nblk(long long n, int t) { return (unsigned)std::max<long long>(1, (n + t - 1) / t); }

static TreeP make_tree(const se3_octree* t) {
    TreeP T;
    T.keys = (const unsigned long long*)t->keys;
    T.n = t->n; T.leaf_size = t->leaf_size; T.max_depth = t->max_depth; T.cap = t->cell_cap;
    T.start = t->cell_start; T.count = t->cell_count; T.level = t->cell_level; T.parent = t->cell_parent;
    T.first_child = t->cell_first_child; T.nchild = t->cell_nchild; T.ckey = (unsigned long long*)t->cell_key;
    T.level_ptr = t->level_ptr;
    T.bsum = nullptr; T.flags = nullptr;
    return T;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int se3_octree_work_bytes(int64_t n, int64_t cell_cap, size_t* bytes) {
    if (n < 0 || !bytes) { set_error("bad argument"); return SE3_ERR_INVALID; }
    const long long nb = (n + RS_TILE - 1) / RS_TILE + 1;
    size_t b = 0;
    b += align256(sizeof(unsigned long long) * (size_t)std::max<int64_t>(n, 1));  // keys alt
    b += align256(sizeof(int) * (size_t)std::max<int64_t>(n, 1));                 // vals alt
    b += align256(sizeof(int) * 256 * (size_t)nb);                                // digit counts
    b += align256(sizeof(int) * 256);                                             // row totals
    b += align256(sizeof(int) * (size_t)(cell_cap / 256 + 2));                    // split block sums
    b += align256(sizeof(int) * 16);                                              // flags
    b += align256(sizeof(unsigned) * 8);                                          // bbox words
    *bytes = b;
    return SE3_OK;
}

extern "C" int se3_octree_build(const float* pos, se3_octree* t, int64_t* m_out, int32_t* nlevels_out, void* stream) {
    if (!pos || !t || !m_out || !nlevels_out) { set_error("null argument"); return SE3_ERR_INVALID; }
    const long long n = t->n;
    if (n < 1 || n > 2000000000ll) { set_error("n must be in [1, 2e9]"); return SE3_ERR_INVALID; }
    if (t->max_depth < 1 || t->max_depth > MAXD || t->leaf_size < 1) { set_error("bad leaf_size / max_depth"); return SE3_ERR_INVALID; }
    if (t->max_depth != MAXD) { set_error("max_depth must be 21 (63-bit keys)"); return SE3_ERR_INVALID; }
    size_t need = 0;
    se3_octree_work_bytes(n, t->cell_cap, &need);
    if (!t->work || t->work_bytes < need) { set_error("workspace too small (%zu < %zu)", t->work_bytes, need); return SE3_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    // carve workspace
    char* w = (char*)t->work;
    const long long nb = (n + RS_TILE - 1) / RS_TILE;
    unsigned long long* kalt = (unsigned long long*)w; w += align256(sizeof(unsigned long long) * (size_t)n);
    int* valt = (int*)w; w += align256(sizeof(int) * (size_t)n);
    int* counts = (int*)w; w += align256(sizeof(int) * 256 * (size_t)(nb + 1));
    int* rowtot = (int*)w; w += align256(sizeof(int) * 256);
    int* bsum = (int*)w; w += align256(sizeof(int) * (size_t)(t->cell_cap / 256 + 2));
    int* flags = (int*)w; w += align256(sizeof(int) * 16);
    unsigned* bb = (unsigned*)w;

    bbox_init_kernel<<<1, 32, 0, st>>>(bb); SE3_LAUNCHED();
    const int bgrid = (int)std::min<long long>(nblk(n, 256), (long long)num_sms() * 8);
    bbox_kernel<<<bgrid, 256, 0, st>>>(pos, n, bb); SE3_LAUNCHED();
    unsigned long long* keys = (unsigned long long*)t->keys;
    morton_kernel<<<nblk(n, 256), 256, 0, st>>>(pos, n, bb, keys, t->order, t->bbox); SE3_LAUNCHED();
    // 8 LSD passes; ping-pong keys/order <-> alt; even number of passes ends in keys/order
    unsigned long long* kin = keys; int* vin = t->order;
    unsigned long long* kout = kalt; int* vout = valt;
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 8 * pass;
        rs_hist_kernel<<<(unsigned)nb, RS_THREADS, 0, st>>>(kin, n, shift, (int)nb, counts); SE3_LAUNCHED();
        rs_scan_rows_kernel<<<256, 256, 0, st>>>(counts, (int)nb, rowtot); SE3_LAUNCHED();
        rs_scatter_kernel<<<(unsigned)nb, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, (int)nb, counts, rowtot); SE3_LAUNCHED();
        std::swap(kin, kout); std::swap(vin, vout);
    }
    // tree
    TreeP T = make_tree(t);
    T.bsum = bsum; T.flags = flags;
    tree_init_kernel<<<1, 32, 0, st>>>(T); SE3_LAUNCHED();
    const long long lvl_cap = std::min<long long>(t->cell_cap, 8ll * (n / (t->leaf_size + 1) + 1) + 8);
    static int coop_blocks = -1;   // resident 256-thread blocks of the all-levels kernel (0: cooperative launch unavailable)
    if (coop_blocks < 0) {
        int dev = 0, coop = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        if (coop && !getenv("SE3_OCTREE_LEVEL_LAUNCHES") &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, split_levels_kernel, 256, 0) == cudaSuccess && per_sm > 0)
            coop_blocks = se3::num_sms() * per_sm;
        else
            coop_blocks = 0;
    }
    // measured: 0.72 -> 0.61 ms at 100k points, equal at 1M, but 5.2 -> 9.2 ms at 10M (the per-level launches with
    // their own grids win once a level holds millions of cells): small clouds only
    if (coop_blocks > 0 && n <= 2000000) {
        const int g = (int)std::min<long long>(coop_blocks, std::max<long long>(1, nblk(lvl_cap, 256)));
        void* args[] = {(void*)&T};
        SE3_CUDA_TRY(cudaLaunchCooperativeKernel((void*)split_levels_kernel, dim3(g), dim3(256), args, 0, st));
        SE3_LAUNCHED();
    } else {
        for (int lev = 0; lev < t->max_depth; ++lev) {
            long long ub = lev < 20 ? std::min<long long>(1ll << (3 * lev), lvl_cap) : lvl_cap;
            const unsigned g = nblk(ub, 256);
            split_count_kernel<<<g, 256, 0, st>>>(T, lev); SE3_LAUNCHED();
            split_scan_kernel<<<1, 256, 0, st>>>(T, lev); SE3_LAUNCHED();
            split_emit_kernel<<<g, 256, 0, st>>>(T, lev); SE3_LAUNCHED();
        }
    }
    int h_lp[MAXD + 2];
    int h_flags[2];
    SE3_CUDA_TRY(cudaMemcpyAsync(h_lp, t->level_ptr, sizeof(int) * (t->max_depth + 2), cudaMemcpyDeviceToHost, st));
    SE3_CUDA_TRY(cudaMemcpyAsync(h_flags, flags, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
    SE3_CUDA_TRY(cudaStreamSynchronize(st));
    if (h_flags[0]) { set_error("cell capacity %lld exceeded at level %d", (long long)t->cell_cap, h_flags[1]); return SE3_ERR_TOO_LARGE; }
    const int m = h_lp[t->max_depth + 1];
    int nlev = 0;
    for (int l = 0; l <= t->max_depth; ++l)
        if (h_lp[l + 1] > h_lp[l]) nlev = l + 1;
    *m_out = m;
    *nlevels_out = nlev;
    leaf_assign_kernel<<<nblk((long long)m * 32, 256), 256, 0, st>>>(T, t->order, t->leaf_of_rank, t->cell_of_particle, m); SE3_LAUNCHED();
    return SE3_OK;
}

extern "C" int se3_graph_degrees(const se3_octree* t, int64_t m, int32_t* nbr, int32_t* deg, int64_t* rowptr,
                                 int64_t* scan_work, int64_t* e_out, void* stream) {
    if (!t || !nbr || !deg || !rowptr || !scan_work || !e_out || m < 1) { set_error("bad argument"); return SE3_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    TreeP T = make_tree(t);
    const long long n = t->n, nn = n + m;
    cell_degree_kernel<<<nblk(m, 128), 128, 0, st>>>(T, (int)m, n, nbr, deg); SE3_LAUNCHED();
    particle_degree_kernel<<<nblk(n, 256), 256, 0, st>>>(T, n, t->leaf_of_rank, deg); SE3_LAUNCHED();
    const long long nb = (nn + 1023) / 1024;
    scan_bsum_kernel<<<(unsigned)nb, 256, 0, st>>>(deg, nn, (long long*)scan_work); SE3_LAUNCHED();
    scan_top_kernel<<<1, 1024, 0, st>>>((long long*)scan_work, nb, (long long*)rowptr + nn); SE3_LAUNCHED();
    scan_final_kernel<<<(unsigned)nb, 256, 0, st>>>(deg, nn, (const long long*)scan_work, (long long*)rowptr); SE3_LAUNCHED();
    long long e = 0;
    SE3_CUDA_TRY(cudaMemcpyAsync(&e, rowptr + nn, sizeof(long long), cudaMemcpyDeviceToHost, st));
    SE3_CUDA_TRY(cudaStreamSynchronize(st));
    *e_out = e;
    return SE3_OK;
}

extern "C" int se3_graph_emit(const se3_octree* t, int64_t m, const int32_t* nbr, const int64_t* rowptr, int32_t* col,
                              int32_t* dst, void* stream) {
    if (!t || !nbr || !rowptr || !col || !dst || m < 1) { set_error("bad argument"); return SE3_ERR_INVALID; }
    TreeP T = make_tree(t);
    emit_edges_kernel<<<nblk((long long)m * 32, 256), 256, 0, (cudaStream_t)stream>>>(T, (int)m, t->n, nbr, (const long long*)rowptr, col, dst);
    SE3_LAUNCHED();
    return SE3_OK;
}

extern "C" int se3_node_data(const se3_octree* t, int64_t m, int32_t nlevels, const float* pos, const float* vel,
                             const float* mass, float* npos, float* nvel, float* nmass, void* stream) {
    if (!t || !pos || !vel || !mass || !npos || !nvel || !nmass || m < 1) { set_error("bad argument"); return SE3_ERR_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    TreeP T = make_tree(t);
    const long long n = t->n;
    permute_particles_kernel<<<nblk(n, 256), 256, 0, st>>>(n, t->order, pos, vel, mass, npos, nvel, nmass); SE3_LAUNCHED();
    const long long lvl_cap = std::min<long long>(t->cell_cap, 8ll * (n / (t->leaf_size + 1) + 1) + 8);
    for (int lev = nlevels - 1; lev >= 0; --lev) {
        long long ub = lev < 20 ? std::min<long long>(1ll << (3 * lev), lvl_cap) : lvl_cap;
        ub = std::min<long long>(ub, m);
        cell_moments_kernel<<<nblk(ub, 128), 128, 0, st>>>(T, lev, n, npos, nvel, nmass); SE3_LAUNCHED();
    }
    return SE3_OK;
}

extern "C" int se3_edge_geometry(int64_t n, int64_t m, int64_t e, const int64_t* rowptr, const int32_t* col,
                                 const int32_t* dst, const float* npos, const float* nvel, const float* nmass,
                                 float mass_scale, float* edge_attr, float* edge_extra, float* node_attr, float* x_in,
                                 void* stream) {
    if (!rowptr || !col || !dst || !npos || !nvel || !nmass || !edge_attr || !edge_extra || !node_attr || !x_in) {
        set_error("null argument");
        return SE3_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (e > 0) { edge_geom_kernel<<<nblk(e, 256), 256, 0, st>>>(e, dst, col, npos, nmass, mass_scale, edge_attr, edge_extra); SE3_LAUNCHED(); }
    node_feat_kernel<<<nblk(n + m, 256), 256, 0, st>>>(n + m, n, (const long long*)rowptr, edge_attr, npos, nvel, nmass, mass_scale, node_attr, x_in);
    SE3_LAUNCHED();
    return SE3_OK;
}

extern "C" int se3_edge_geometry_l2(int64_t n, int64_t m, int64_t e, const int64_t* rowptr, const int32_t* col,
                                    const int32_t* dst, const float* npos, const float* nvel, float* edge_attr9,
                                    float* node_attr9, void* stream) {
    if (!rowptr || !col || !dst || !npos || !nvel || !edge_attr9 || !node_attr9) {
        set_error("null argument");
        return SE3_ERR_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (e > 0) { edge_sh2_kernel<<<nblk(e, 256), 256, 0, st>>>(e, dst, col, npos, edge_attr9); SE3_LAUNCHED(); }
    node_sh2_kernel<<<nblk(n + m, 256), 256, 0, st>>>(n + m, (const long long*)rowptr, edge_attr9, nvel, node_attr9);
    SE3_LAUNCHED();
    return SE3_OK;
}
