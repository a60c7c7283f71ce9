// Tile programs of the l <= 2 tensor product (forward and backward), written once and compiled twice: by nvcc inside
// o3tp.cu (one CUDA thread per `tid`) and by g++ inside tests/emu/o3tp_emu.cpp (threads run one after another, phase by
// phase) so the table walking and indexing are checked against the oracle on a machine without a GPU.
//
// The includer defines (and includes o3tp_cg_gen.inl first):
//   O3_DEV                     function qualifier
//   O3_THREADS / O3_END        open / close a region every thread of the block executes; O3_END is a block barrier
//   O3_ATOMIC_ADD(p, v)        shared-memory float add that tolerates several threads on one address
//   O3_GW_ADD(S, p, v)         weight-gradient accumulation: plain add, or a global atomic add when S.gw_global
//   O3_MULHI(a, b)             high 32 bits of the unsigned 32 x 32 product
//   O3_CP4(dst, src) / O3_CP_COMMIT() / O3_CP_WAIT()   4-byte asynchronous global -> shared copy, group commit, wait all
//   O3_ACC_DECL / O3_ACC(acc, slot, tid)   per-thread accumulator sets [MAXIO_GW][16] that live across regions and
//                              tiles (registers under nvcc, one array per emulated thread otherwise)
//   O3_GLOBAL_ADD(p, v) / O3_GLOBAL_ADD4(p, a, b, c, d)   atomic float add to global memory (one / four consecutive)
//   O3_I2F(i)                  reinterpret an int32 table word as float
//   O3_NT_DECL                 extra parameter `, int NT_` carrying the emulated block size (empty under nvcc)
//   o3f4 / O3_LD4(p)           four consecutive floats read from a 16-byte aligned shared-memory address
//   O3_UNROLL / O3_UNROLL2     unroll pragmas (full / by two)
//
// Math (oracle/lmax2_oracle.py forward): for every output irrep io with stacked paths,
//   f[kk, e, c]   = sum_{i,j} C_p[i,j,c] x1[e, off1_p + u d1 + i] y[e, off2_p + j]     kk = koff_p + u
//   out[e, w, c]  = a_io sum_kk f[kk, e, c] W_io[kk, w]
// backward, with g' = a_io g:
//   gW_io[kk, w]  = sum_{e,c} f[kk, e, c] g'[e, w, c]
//   G[kk, e, c]   = sum_w W_io[kk, w] g'[e, w, c]
//   gx1[e, off1 + u d1 + i] += sum C[i,j,c] y[e, j] G[koff + u, e, c];   gy[e, j] += sum C[i,j,c] x1[...] G[...]
// Per lane (= row e) and path the second input is folded once into M[i][c] = sum_j C[i][j][c] y[j], so that
// f[c] = sum_i M[i][c] x[i] and gx[i] = sum_c M[i][c] G[c].

// in1 as a virtual concatenation of up to 4 row segments (se3_rowseg of the C ABI): row r, columns [c0, c0 + width) of
// segment s = base[s][(idx[s] ? idx[s][r] : r) * ld[s] + 0 .. width).  The gradient goes back per segment: stored
// (identity rows), added atomically (gathered rows) or skipped.  Loops over the segments are fully unrolled so that the
// kernel-parameter arrays are only indexed statically.
struct O3Rows {
    const float* base[4];
    const int* idx[4];
    int ld[4], c0[4], width[4];
    int nseg;
};
struct O3GRows {
    float* base[4];
    const int* idx[4];
    int ld[4], c0[4], width[4], mode[4];  // see o3_tile_store_grad
    int nseg;
};

O3_DEV void o3_row_load_async(const O3Rows& X, long long row, float* dst, int lane) {
    O3_UNROLL
    for (int s = 0; s < 4; ++s)
        if (s < X.nseg) {
            const float* src = X.base[s] + (X.idx[s] ? (long long)X.idx[s][row] : row) * X.ld[s];
            for (int c = lane; c < X.width[s]; c += 32) O3_CP4(dst + X.c0[s] + c, src + c);
        }
}
O3_DEV void o3_row_load(const O3Rows& X, long long row, float* dst, int lane) {
    O3_UNROLL
    for (int s = 0; s < 4; ++s)
        if (s < X.nseg) {
            const float* src = X.base[s] + (X.idx[s] ? (long long)X.idx[s][row] : row) * X.ld[s];
            for (int c = lane; c < X.width[s]; c += 32) dst[X.c0[s] + c] = src[c];
        }
}
// Gradient tile [nrow, D1p] (shared) -> the segments' destinations.  mode & 15: 1 store (identity rows), 2 atomic add per
// row, 3 rows sorted by index: the rows of a run inside the tile are summed first and added once (an octree node has
// ~18 incoming edges, so this removes most of the atomics of the dst segment); mode & 16: 16-byte reductions (width, ld
// multiples of 4, base 16-byte aligned: checked by the host).
O3_DEV void o3_tile_store_grad(const O3GRows& G, long long row0, int nrow, const float* gxs, int D1p, int tid, int NT) {
    const int warp = tid >> 5, lane = tid & 31, nw = NT >> 5;
    O3_UNROLL
    for (int s = 0; s < 4; ++s)
        if (s < G.nseg && G.mode[s] != 0) {
            const int mode = G.mode[s] & 15, c0 = G.c0[s], width = G.width[s];
            const bool v4 = (G.mode[s] & 16) != 0;
            for (int e = warp; e < nrow; e += nw) {
                const float* src = gxs + e * D1p + c0;
                if (mode == 1) {
                    float* dst = G.base[s] + (row0 + e) * G.ld[s];
                    for (int c = lane; c < width; c += 32) dst[c] = src[c];
                    continue;
                }
                const int id = G.idx[s][row0 + e];
                int len = 1;
                if (mode == 3) {
                    if (e > 0 && G.idx[s][row0 + e - 1] == id) continue;   // not the head of its run (warp-uniform)
                    while (e + len < nrow && G.idx[s][row0 + e + len] == id) ++len;
                }
                float* dst = G.base[s] + (long long)id * G.ld[s];
                if (v4) {
                    for (int c = 4 * lane; c < width; c += 128) {
                        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                        for (int k = 0; k < len; ++k) {
                            const float* p = src + k * D1p + c;
                            a0 += p[0]; a1 += p[1]; a2 += p[2]; a3 += p[3];
                        }
                        O3_GLOBAL_ADD4(dst + c, a0, a1, a2, a3);
                    }
                } else {
                    for (int c = lane; c < width; c += 32) {
                        float a0 = 0.f;
                        for (int k = 0; k < len; ++k) a0 += src[k * D1p + c];
                        O3_GLOBAL_ADD(dst + c, a0);
                    }
                }
            }
        }
}

struct O3Fwd {
    const int32_t* tab;  // table blob (shared memory)
    const float* Ws;     // all weights, per io [K, IO_MULP] zero padded (shared, resident)
    float *xs0, *xs1, *ys0, *ys1, *os;  // double-buffered input tiles, output tile
    int TE;
};

struct O3Bwd {
    const int32_t* tab;
    const float* WT;  // all weights transposed, per io [mul, 4 * IO_NBLK] in block order (shared, resident)
    float* gWs;       // weight gradient accumulators, flat like the weights (shared, resident)
    float *xs0, *xs1, *ys0, *ys1, *gs0, *gs1;  // double-buffered x / y / cotangent tiles
    float *gxs, *gys, *F, *GT, *scr;
    int gw_global;  // gWs is the global result (weights too large to keep accumulators resident): add atomically
};

// ---- forward: one warp per work unit (output irrep, chunk of CW output channels, group of 32 rows), lane = row.
// Per path the lane folds its spherical-harmonics values into M[i][c] = sum_j C[i][j][c] y[j] once (generated code,
// o3tp_cg_gen.inl), then per input channel u: d1 shared loads, f[c] = sum_i M[i][c] x[i], one broadcast weight row,
// acc[c][t] += f[c] w[t].  No feature buffer, no cross-lane traffic.
template <int L1, int L2, int LO, int CW>
O3_DEV void o3_fwd_path(const float* xr, const float* yr, int mul1, const float* w, int mulp,
                        float (&acc)[2 * LO + 1][CW]) {
    constexpr int D1 = 2 * L1 + 1, DO = 2 * LO + 1;
    float M[D1][DO];
    o3_M<L1, L2, LO>(yr, M);
    for (int u = 0; u < mul1; ++u, xr += D1, w += mulp) {
        float f[DO];
        O3_UNROLL
        for (int c = 0; c < DO; ++c) f[c] = 0.f;
        O3_UNROLL
        for (int i = 0; i < D1; ++i) {
            const float x = xr[i];
            O3_UNROLL
            for (int c = 0; c < DO; ++c)
                if ((o3_nz<L1, L2, LO>::mask >> (i * DO + c)) & 1u) f[c] += M[i][c] * x;
        }
        O3_UNROLL
        for (int t4 = 0; t4 < CW; t4 += 4) {
            const o3f4 wv = O3_LD4(w + t4);
            O3_UNROLL
            for (int c = 0; c < DO; ++c) {
                acc[c][t4] += f[c] * wv.x; acc[c][t4 + 1] += f[c] * wv.y;
                acc[c][t4 + 2] += f[c] * wv.z; acc[c][t4 + 3] += f[c] * wv.w;
            }
        }
    }
}

template <int LO, int CW>
O3_DEV void o3_fwd_unit(const int32_t* tab, const int32_t* IO, const float* xe, const float* ye, const float* Ws,
                        float* oe, int q) {
    constexpr int DO = 2 * LO + 1;
    float acc[DO][CW];
    O3_UNROLL
    for (int c = 0; c < DO; ++c)
        O3_UNROLL
        for (int t = 0; t < CW; ++t) acc[c][t] = 0.f;
    const int mulp = IO[o3::IO_MULP], mul = IO[o3::IO_MUL];
    const float* wq = Ws + IO[o3::IO_WSOFF] + q * CW;
    for (int p = IO[o3::IO_PBEG]; p < IO[o3::IO_PEND]; ++p) {
        const int32_t* P = tab + tab[o3::H_PATH] + p * o3::PATH_W;
        const float* xr = xe + P[o3::P_OFF1];
        const float* yr = ye + P[o3::P_OFF2];
        const float* w = wq + P[o3::P_KOFF] * mulp;
        const int mul1 = P[o3::P_MUL1];
        switch (P[o3::P_L1] * 9 + P[o3::P_L2] * 3 + LO) {
#define O3_CASE(A, B, C)                                                                      \
    case A * 9 + B * 3 + C:                                                                   \
        if constexpr (C == LO) o3_fwd_path<A, B, C, CW>(xr, yr, mul1, w, mulp, acc);          \
        break;
            O3_TRIPLES(O3_CASE)
#undef O3_CASE
        }
    }
    const float a = O3_I2F(IO[o3::IO_A]);
    float* o = oe + IO[o3::IO_OFF] + q * CW * DO;
    O3_UNROLL
    for (int t = 0; t < CW; ++t)
        if (q * CW + t < mul) {
            O3_UNROLL
            for (int c = 0; c < DO; ++c) o[t * DO + c] = a * acc[c][t];
        }
}

// asynchronous copy of one tile of both inputs into buffer `buf` (rows past the end are zero filled)
O3_DEV void o3_fwd_load(const O3Fwd& S, int buf, const O3Rows& in1, const float* __restrict__ in2,
                        long long row0, int nrow, int tid, int NT) {
    const int32_t* tab = S.tab;
    const int D1 = tab[o3::H_D1], D2 = tab[o3::H_D2], D1p = D1 | 1, D2p = D2 | 1;
    float* xs = buf ? S.xs1 : S.xs0;
    float* ys = buf ? S.ys1 : S.ys0;
    const int warp = tid >> 5, lane = tid & 31, nw = NT >> 5;
    for (int e = warp; e < S.TE; e += nw) {
        if (e < nrow) {
            o3_row_load_async(in1, row0 + e, xs + e * D1p, lane);
            if (lane < D2) O3_CP4(ys + e * D2p + lane, in2 + (row0 + e) * D2 + lane);
        } else {
            for (int c = lane; c < D1; c += 32) xs[e * D1p + c] = 0.f;
            if (lane < D2) ys[e * D2p + lane] = 0.f;
        }
    }
    O3_CP_COMMIT();
}

// One tile: the inputs of this tile were requested earlier into buffer `buf` (by the previous call, or by the
// prologue for the first tile); the next tile (nrow_next > 0) is requested into the other buffer before computing.
O3_DEV void o3_fwd_tile(const O3Fwd& S, int buf, const O3Rows& in1, const float* __restrict__ in2,
                        float* __restrict__ out, long long row0, int nrow, long long row0_next,
                        int nrow_next O3_NT_DECL) {
    const int32_t* tab = S.tab;
    const int D1 = tab[o3::H_D1], D2 = tab[o3::H_D2], DO = tab[o3::H_DOUT];
    const int D1p = D1 | 1, D2p = D2 | 1, DOp = DO | 1;
    const float* xs = buf ? S.xs1 : S.xs0;
    const float* ys = buf ? S.ys1 : S.ys0;

    O3_THREADS
        (void)tid; (void)NT;
        O3_CP_WAIT();
    O3_END

    O3_THREADS
        if (nrow_next > 0) o3_fwd_load(S, buf ^ 1, in1, in2, row0_next, nrow_next, tid, NT);
        const int warp = tid >> 5, lane = tid & 31;
        const int32_t* U = tab + tab[o3::H_UNIT];
        for (int k = U[warp]; k < U[warp + 1]; ++k) {
            const int packed = U[o3::NWARP + 1 + k];
            const int io = packed & 255, q = (packed >> 8) & 255, e = (packed >> 16) * 32 + lane;
            const int32_t* IO = tab + tab[o3::H_IO] + io * o3::IO_W;
            const float* xe = xs + e * D1p;
            const float* ye = ys + e * D2p;
            float* oe = S.os + e * DOp;
            switch (IO[o3::IO_D] * 16 + IO[o3::IO_CW]) {
                case 1 * 16 + 4: o3_fwd_unit<0, 4>(tab, IO, xe, ye, S.Ws, oe, q); break;
                case 1 * 16 + 8: o3_fwd_unit<0, 8>(tab, IO, xe, ye, S.Ws, oe, q); break;
                case 1 * 16 + 12: o3_fwd_unit<0, 12>(tab, IO, xe, ye, S.Ws, oe, q); break;
                case 3 * 16 + 4: o3_fwd_unit<1, 4>(tab, IO, xe, ye, S.Ws, oe, q); break;
                case 3 * 16 + 8: o3_fwd_unit<1, 8>(tab, IO, xe, ye, S.Ws, oe, q); break;
                case 3 * 16 + 12: o3_fwd_unit<1, 12>(tab, IO, xe, ye, S.Ws, oe, q); break;
                case 5 * 16 + 4: o3_fwd_unit<2, 4>(tab, IO, xe, ye, S.Ws, oe, q); break;
                case 5 * 16 + 8: o3_fwd_unit<2, 8>(tab, IO, xe, ye, S.Ws, oe, q); break;
            }
        }
    O3_END

    O3_THREADS
        const int warp = tid >> 5, lane = tid & 31, nw = NT >> 5;
        for (int e = warp; e < nrow; e += nw) {
            float* dst = out + (row0 + e) * DO;
            for (int c = lane; c < DO; c += 32) dst[c] = S.os[e * DOp + c];
        }
    O3_END
}

// ---- backward.  Tile of 32 rows, lane = row.  Per output irrep: GT[w][e*d + c] = a g[e][w][c] (shared), then rounds of
// NWARP blocks (<= 4 channels of one input irrep, all its paths into this output; one block per warp):
//   step 1 (warp = block), per path: Gc[uu][c] = sum_w W[kk0+uu][w] GT[w][e, c] from one transposed weight float4 per w;
//           per channel f -> F rows of the warp's slot, gx accumulated in registers over the paths (then added to the
//           shared gx tile, which this block owns), P[i][c] += x[i] Gc[uu][c] -> gy[j] += C . P (generated code).
//   step 2 (all threads): weight gradient of the round's sub-blocks (4 feature rows of one path), 4 x 4 register blocks
//           of F . GT^T; the (e, c) axis is cut into slices so that every thread has work, the slices' partial sums
//           go through a scratch array and are added to the resident accumulators at the start of the next region.
template <int A, int B, int C> struct o3_tri { static constexpr bool v = C >= (A > B ? A - B : B - A) && C <= A + B; };

template <int L1, int L2, int LO>
O3_DEV void o3_bwd_path(const float* xr, const float* yr, float (&gx)[4][2 * L1 + 1], float* gyr, const float* GTe,
                        int Rp, const float* wt, int KPP, int mul, int nu, float* Fe) {
    constexpr int D1 = 2 * L1 + 1, D2 = 2 * L2 + 1, DO = 2 * LO + 1;
    constexpr unsigned NZ = o3_nz<L1, L2, LO>::mask;
    float M[D1][DO];
    o3_M<L1, L2, LO>(yr, M);
    float Gc[4][DO];
    O3_UNROLL
    for (int uu = 0; uu < 4; ++uu)
        O3_UNROLL
        for (int c = 0; c < DO; ++c) Gc[uu][c] = 0.f;
    for (int w = 0; w < mul; ++w) {
        const o3f4 t = O3_LD4(wt + w * KPP);
        O3_UNROLL
        for (int c = 0; c < DO; ++c) {
            const float gv = GTe[(size_t)w * Rp + c];
            Gc[0][c] += t.x * gv; Gc[1][c] += t.y * gv; Gc[2][c] += t.z * gv; Gc[3][c] += t.w * gv;
        }
    }
    float Pm[D1][DO];
    O3_UNROLL
    for (int i = 0; i < D1; ++i)
        O3_UNROLL
        for (int c = 0; c < DO; ++c) Pm[i][c] = 0.f;
    O3_UNROLL
    for (int uu = 0; uu < 4; ++uu) {
        float* f = Fe + (size_t)uu * Rp;
        if (uu < nu) {
            float x[D1], fc[DO];
            O3_UNROLL
            for (int i = 0; i < D1; ++i) x[i] = xr[uu * D1 + i];
            O3_UNROLL
            for (int c = 0; c < DO; ++c) fc[c] = 0.f;
            O3_UNROLL
            for (int i = 0; i < D1; ++i) {
                O3_UNROLL
                for (int c = 0; c < DO; ++c)
                    if ((NZ >> (i * DO + c)) & 1u) {
                        fc[c] += M[i][c] * x[i];
                        gx[uu][i] += M[i][c] * Gc[uu][c];
                        Pm[i][c] += x[i] * Gc[uu][c];
                    }
            }
            O3_UNROLL
            for (int c = 0; c < DO; ++c) f[c] = fc[c];
        } else {
            O3_UNROLL
            for (int c = 0; c < DO; ++c) f[c] = 0.f;
        }
    }
    if (gyr != nullptr) {
        float gy[D2];
        O3_UNROLL
        for (int j = 0; j < D2; ++j) gy[j] = 0.f;
        o3_gy<L1, L2, LO>(Pm, gy);
        O3_UNROLL
        for (int j = 0; j < D2; ++j) O3_ATOMIC_ADD(gyr + j, gy[j]);
    }
}

template <int L1, int LO>
O3_DEV void o3_bwd_group(const int32_t* tab, const int32_t* G, int u0, const float* xe, const float* ye, float* gxe,
                         float* gye, const float* GTe, int Rp, const float* wt0, int KPP, int mul, float* Fe) {
    constexpr int D1 = 2 * L1 + 1;
    const int nu = G[o3::G_MUL1] - u0 < 4 ? G[o3::G_MUL1] - u0 : 4;
    const float* xr = xe + G[o3::G_OFF1] + u0 * D1;
    float gx[4][D1];
    O3_UNROLL
    for (int uu = 0; uu < 4; ++uu)
        O3_UNROLL
        for (int i = 0; i < D1; ++i) gx[uu][i] = 0.f;
    for (int pi = 0; pi < G[o3::G_NP]; ++pi) {
        const int32_t* P = tab + tab[o3::H_PATH] + G[o3::G_P0 + pi] * o3::PATH_W;
        const float* yr = ye + P[o3::P_OFF2];
        float* gyr = gye != nullptr ? gye + P[o3::P_OFF2] : nullptr;
        const float* wt = wt0 + 4 * pi;
        float* Fp = Fe + (size_t)(4 * pi) * Rp;
        switch (P[o3::P_L2]) {
            case 0:
                if constexpr (o3_tri<L1, 0, LO>::v) o3_bwd_path<L1, 0, LO>(xr, yr, gx, gyr, GTe, Rp, wt, KPP, mul, nu, Fp);
                break;
            case 1:
                if constexpr (o3_tri<L1, 1, LO>::v) o3_bwd_path<L1, 1, LO>(xr, yr, gx, gyr, GTe, Rp, wt, KPP, mul, nu, Fp);
                break;
            case 2:
                if constexpr (o3_tri<L1, 2, LO>::v) o3_bwd_path<L1, 2, LO>(xr, yr, gx, gyr, GTe, Rp, wt, KPP, mul, nu, Fp);
                break;
        }
    }
    float* gxr = gxe + G[o3::G_OFF1] + u0 * D1;
    O3_UNROLL
    for (int uu = 0; uu < 4; ++uu)
        if (uu < nu) {
            O3_UNROLL
            for (int i = 0; i < D1; ++i) gxr[uu * D1 + i] += gx[uu][i];
        }
}

// pending partial sums of the previous step 2 -> resident weight-gradient accumulators
struct O3Pending {
    const int32_t* IO;
    int sbeg, nsb, lg;  // lg = log2(slices); -1: nothing pending
};

// Scratch of the sliced partial sums: scr[k * O3_SCR_LD + (t << lg) + s] for entry k of the 4 x 4 block of item t, slice
// s.  Writers (lanes = consecutive items) and readers (lanes = consecutive k, 16-byte loads over s) are conflict free.
#define O3_SCR_LD (32 * o3::NWARP + 4)

O3_DEV void o3_bwd_reduce(const O3Bwd& S, const O3Pending& Q, int tid, int NT) {
    if (Q.lg <= 0) return;
    const int32_t* tab = S.tab;
    const int32_t* IO = Q.IO;
    const int mul = IO[o3::IO_MUL], nwb = ((mul + 3) & ~3) >> 2, base = Q.nsb * nwb;
    const unsigned magic = (unsigned)IO[o3::IO_NWB_MAGIC];
    const int32_t* BL = tab + tab[o3::H_BLK] + IO[o3::IO_BLK];
    const int32_t* SUB = tab + tab[o3::H_SUB] + IO[o3::IO_SUB];
    for (int o = tid; o < (base << 4); o += NT) {     // one of the 16 entries of a 4 x 4 block per thread
        const int k = o & 15, t = o >> 4;
        float sum = 0.f;
        const float* ps = S.scr + k * O3_SCR_LD + (t << Q.lg);
        if (Q.lg >= 2) {
            for (int s4 = 0; s4 < (1 << Q.lg); s4 += 4) {
                const o3f4 v = O3_LD4(ps + s4);
                sum += (v.x + v.y) + (v.z + v.w);
            }
        } else {
            sum = ps[0] + ps[1];
        }
        const int sb = nwb == 1 ? t : (int)O3_MULHI((unsigned)t, magic), wb = t - sb * nwb;
        const int word = SUB[Q.sbeg + sb];
        const int32_t* B = BL + (word & 0xffff) * o3::BLK_W;
        const int32_t* G = tab + tab[o3::H_GRP] + (B[o3::B_GRP] & 0xffff) * o3::GRP_W;
        const int32_t* P = tab + tab[o3::H_PATH] + G[o3::G_P0 + (word >> 16)] * o3::PATH_W;
        const int u = (B[o3::B_GRP] >> 16) + (k >> 2), w = 4 * wb + (k & 3);
        if (u < G[o3::G_MUL1] && w < mul) O3_GW_ADD(S, S.gWs + P[o3::P_WOFF] + u * mul + w, sum);
    }
}

// asynchronous copy of one tile of both inputs and of the cotangent into buffer `buf` (rows past the end: zeros)
O3_DEV void o3_bwd_load(const O3Bwd& S, int buf, const O3Rows& in1, const float* __restrict__ in2,
                        const float* __restrict__ gout, long long row0, int nrow, int tid, int NT) {
    const int32_t* tab = S.tab;
    const int D1 = tab[o3::H_D1], D2 = tab[o3::H_D2], DO = tab[o3::H_DOUT];
    const int D1p = D1 | 1, D2p = D2 | 1, DOp = DO | 1;
    float* xs = buf ? S.xs1 : S.xs0;
    float* ys = buf ? S.ys1 : S.ys0;
    float* gs = buf ? S.gs1 : S.gs0;
    const int warp = tid >> 5, lane = tid & 31, nw = NT >> 5;
    for (int e = warp; e < o3::TE_BWD; e += nw) {
        if (e < nrow) {
            o3_row_load_async(in1, row0 + e, xs + e * D1p, lane);
            if (lane < D2) O3_CP4(ys + e * D2p + lane, in2 + (row0 + e) * D2 + lane);
            const float* gsrc = gout + (row0 + e) * DO;
            for (int c = lane; c < DO; c += 32) O3_CP4(gs + e * DOp + c, gsrc + c);
        } else {
            for (int c = lane; c < D1; c += 32) xs[e * D1p + c] = 0.f;
            if (lane < D2) ys[e * D2p + lane] = 0.f;
            for (int c = lane; c < DO; c += 32) gs[e * DOp + c] = 0.f;
        }
    }
    O3_CP_COMMIT();
}

// One tile; its inputs were requested earlier into buffer `buf`, the next tile (nrow_next > 0) is requested into the
// other buffer at the start of the first compute region.
O3_DEV void o3_bwd_tile(const O3Bwd& S, int buf, const O3Rows& in1, const float* __restrict__ in2,
                        const float* __restrict__ gout, const O3GRows& gin1, float* __restrict__ gin2,
                        long long row0, int nrow, long long row0_next, int nrow_next O3_NT_DECL) {
    const int32_t* tab = S.tab;
    const int D1 = tab[o3::H_D1], D2 = tab[o3::H_D2], DO = tab[o3::H_DOUT], nio = tab[o3::H_NIO];
    const int D1p = D1 | 1, D2p = D2 | 1, DOp = DO | 1;
    constexpr int TE = o3::TE_BWD;
    const float* xs = buf ? S.xs1 : S.xs0;
    const float* ys = buf ? S.ys1 : S.ys0;
    const float* gs = buf ? S.gs1 : S.gs0;
    O3Pending Q;
    Q.IO = nullptr; Q.sbeg = 0; Q.nsb = 0; Q.lg = -1;

    O3_THREADS
        O3_CP_WAIT();
        const int warp = tid >> 5, lane = tid & 31, nw = NT >> 5;
        for (int e = warp; e < TE; e += nw) {
            for (int c = lane; c < D1; c += 32) S.gxs[e * D1p + c] = 0.f;
            if (lane < D2) S.gys[e * D2p + lane] = 0.f;
        }
    O3_END
    bool prefetch = nrow_next > 0;

    for (int io = 0; io < nio; ++io) {
        const int32_t* IO = tab + tab[o3::H_IO] + io * o3::IO_W;
        const int mul = IO[o3::IO_MUL], d = IO[o3::IO_D], nblk = IO[o3::IO_NBLK], nsub = IO[o3::IO_NSUB];
        const int mulp = (mul + 3) & ~3, nwb = mulp >> 2, R = TE * d, Rp = R | 1, KPP = 4 * nsub;
        const int32_t* BL = tab + tab[o3::H_BLK] + IO[o3::IO_BLK];
        const int32_t* SUB = tab + tab[o3::H_SUB] + IO[o3::IO_SUB];
        if (nblk == 0) continue;  // block-uniform
        O3_THREADS
            if (prefetch) o3_bwd_load(S, buf ^ 1, in1, in2, gout, row0_next, nrow_next, tid, NT);
            o3_bwd_reduce(S, Q, tid, NT);
            const float a = O3_I2F(IO[o3::IO_A]);
            const int warp = tid >> 5, lane = tid & 31, nw = NT >> 5;
            for (int w = warp; w < mulp; w += nw)
                for (int c = 0; c < d; ++c)
                    S.GT[(size_t)w * Rp + lane * d + c] = w < mul ? a * gs[lane * DOp + IO[o3::IO_OFF] + w * d + c] : 0.f;
        O3_END
        Q.lg = -1;
        prefetch = false;
        for (int b0 = 0; b0 < nblk; b0 += o3::NWARP) {
            O3_THREADS
                o3_bwd_reduce(S, Q, tid, NT);
                const int warp = tid >> 5, e = tid & 31, b = b0 + warp;
                if (b < nblk) {
                    const int32_t* B = BL + b * o3::BLK_W;
                    const int32_t* G = tab + tab[o3::H_GRP] + (B[o3::B_GRP] & 0xffff) * o3::GRP_W;
                    const int u0 = B[o3::B_GRP] >> 16;
                    const float* xe = xs + e * D1p;
                    const float* ye = ys + e * D2p;
                    float* gxe = S.gxs + e * D1p;
                    float* gye = gin2 != nullptr ? S.gys + e * D2p : nullptr;
                    const float* GTe = S.GT + e * d;
                    const float* wt0 = S.WT + IO[o3::IO_WTOFF] + 4 * B[o3::B_SUB0];
                    float* Fe = S.F + (size_t)(4 * tab[o3::H_MAXNP] * warp) * Rp + e * d;
                    switch (G[o3::G_L1] * 3 + (d >> 1)) {
#define O3_CASE(A, C)                                                                            \
    case A * 3 + C:                                                                              \
        o3_bwd_group<A, C>(tab, G, u0, xe, ye, gxe, gye, GTe, Rp, wt0, KPP, mul, Fe);            \
        break;
                        O3_CASE(0, 0) O3_CASE(0, 1) O3_CASE(0, 2) O3_CASE(1, 0) O3_CASE(1, 1) O3_CASE(1, 2)
                        O3_CASE(2, 0) O3_CASE(2, 1) O3_CASE(2, 2)
#undef O3_CASE
                    }
                }
            O3_END
            const int nb = nblk - b0 < o3::NWARP ? nblk - b0 : o3::NWARP;
            const int sbeg = BL[b0 * o3::BLK_W + o3::B_SUB0];
            const int send = b0 + nb < nblk ? BL[(b0 + nb) * o3::BLK_W + o3::B_SUB0] : nsub;
            Q.IO = IO; Q.sbeg = sbeg; Q.nsb = send - sbeg;
            {   // slices: the largest power of two that keeps one item per thread
                const int base = Q.nsb * nwb;
                Q.lg = 0;
                while (Q.lg < 5 && (base << (Q.lg + 1)) <= (int)(32 * o3::NWARP)) ++Q.lg;
            }
            O3_THREADS
                const int lg = Q.lg, ns = 1 << lg;
                const unsigned magic = (unsigned)IO[o3::IO_NWB_MAGIC];
                for (int item = tid; item < ((Q.nsb * nwb) << lg); item += NT) {
                    const int s = item & (ns - 1), t = item >> lg;
                    const int sb = nwb == 1 ? t : (int)O3_MULHI((unsigned)t, magic), wb = t - sb * nwb;
                    const int word = SUB[sbeg + sb], bl = (word & 0xffff) - b0, pi = word >> 16;
                    const float* f = S.F + (size_t)(4 * (tab[o3::H_MAXNP] * bl + pi)) * Rp;
                    const float* g = S.GT + (size_t)(4 * wb) * Rp;
                    const int r0 = (s * R) >> lg, r1 = ((s + 1) * R) >> lg;
                    float acc[4][4];
                    O3_UNROLL
                    for (int i = 0; i < 4; ++i)
                        O3_UNROLL
                        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
                    O3_UNROLL2
                    for (int r = r0; r < r1; ++r) {
                        const float f0 = f[r], f1 = f[Rp + r], f2 = f[2 * Rp + r], f3 = f[3 * Rp + r];
                        const float g0 = g[r], g1 = g[Rp + r], g2 = g[2 * Rp + r], g3 = g[3 * Rp + r];
                        acc[0][0] += f0 * g0; acc[0][1] += f0 * g1; acc[0][2] += f0 * g2; acc[0][3] += f0 * g3;
                        acc[1][0] += f1 * g0; acc[1][1] += f1 * g1; acc[1][2] += f1 * g2; acc[1][3] += f1 * g3;
                        acc[2][0] += f2 * g0; acc[2][1] += f2 * g1; acc[2][2] += f2 * g2; acc[2][3] += f2 * g3;
                        acc[3][0] += f3 * g0; acc[3][1] += f3 * g1; acc[3][2] += f3 * g2; acc[3][3] += f3 * g3;
                    }
                    if (lg > 0) {   // item < NT here: one scratch column per thread
                        O3_UNROLL
                        for (int i = 0; i < 4; ++i)
                            O3_UNROLL
                            for (int j = 0; j < 4; ++j) S.scr[(4 * i + j) * O3_SCR_LD + item] = acc[i][j];
                    } else {        // this thread owns the 4 x 4 weight block
                        const int32_t* B = BL + (word & 0xffff) * o3::BLK_W;
                        const int32_t* G = tab + tab[o3::H_GRP] + (B[o3::B_GRP] & 0xffff) * o3::GRP_W;
                        const int32_t* P = tab + tab[o3::H_PATH] + G[o3::G_P0 + pi] * o3::PATH_W;
                        const int u0 = B[o3::B_GRP] >> 16;
                        float* gw = S.gWs + P[o3::P_WOFF] + u0 * mul + 4 * wb;
                        O3_UNROLL
                        for (int i = 0; i < 4; ++i)
                            O3_UNROLL
                            for (int j = 0; j < 4; ++j)
                                if (u0 + i < G[o3::G_MUL1] && 4 * wb + j < mul) O3_GW_ADD(S, gw + i * mul + j, acc[i][j]);
                    }
                }
            O3_END
        }
    }

    O3_THREADS
        if (prefetch) o3_bwd_load(S, buf ^ 1, in1, in2, gout, row0_next, nrow_next, tid, NT);  // no output irrep had paths
        o3_bwd_reduce(S, Q, tid, NT);
        o3_tile_store_grad(gin1, row0, nrow, S.gxs, D1p, tid, NT);
        const int warp = tid >> 5, lane = tid & 31, nw = NT >> 5;
        if (gin2 != nullptr)
            for (int e = warp; e < nrow; e += nw)
                if (lane < D2) gin2[(row0 + e) * D2 + lane] = S.gys[e * D2p + lane];
    O3_END
}

// ======================================================================================================================
// Split backward (used when the plan allows it, H_SPLIT): two kernels with small shared-memory footprints instead of
// the fused one above (which stays as the general fallback).
//
// (1) input gradients: forward-like.  One warp per unit (in1 irrep, block of 4 channels, 32-row group), lane = row.
//     For every output irrep the block reaches and every path: G[uu][c] = sum_w aW^T[w][kk0+uu] g[e][w][c] straight from
//     the cotangent tile (row stride odd: conflict free), gx[uu][i] += sum_c M[i][c] G[uu][c] in registers, then one
//     exclusive store into the shared gx tile.  No feature buffer, no scratch, no atomics (except the optional in2
//     gradient).  The staged transposed weights carry the normalisation factor a.
struct O3Gin {
    const int32_t* tab;
    const float* WT;  // a * W^T per io [mul, 4 * IO_NSUB] (shared, resident)
    float *xs, *ys, *gs, *gxs, *gys;
    int need_gy;
};

template <int L1, int L2, int LO>
O3_DEV void o3_gin_path(const float* xr, const float* yr, float (&gx)[4][2 * L1 + 1], float* gyr, const float* ge,
                        const float* wt, int KPP, int mul, int nu) {
    constexpr int D1 = 2 * L1 + 1, D2 = 2 * L2 + 1, DO = 2 * LO + 1;
    constexpr unsigned NZ = o3_nz<L1, L2, LO>::mask;
    float M[D1][DO];
    o3_M<L1, L2, LO>(yr, M);
    float Gc[4][DO];
    O3_UNROLL
    for (int uu = 0; uu < 4; ++uu)
        O3_UNROLL
        for (int c = 0; c < DO; ++c) Gc[uu][c] = 0.f;
    for (int w = 0; w < mul; ++w) {
        const o3f4 t = O3_LD4(wt + w * KPP);
        O3_UNROLL
        for (int c = 0; c < DO; ++c) {
            const float gv = ge[w * DO + c];
            Gc[0][c] += t.x * gv; Gc[1][c] += t.y * gv; Gc[2][c] += t.z * gv; Gc[3][c] += t.w * gv;
        }
    }
    O3_UNROLL
    for (int uu = 0; uu < 4; ++uu)
        O3_UNROLL
        for (int i = 0; i < D1; ++i)
            O3_UNROLL
            for (int c = 0; c < DO; ++c)
                if ((NZ >> (i * DO + c)) & 1u) gx[uu][i] += M[i][c] * Gc[uu][c];
    if (gyr != nullptr) {
        float Pm[D1][DO];
        O3_UNROLL
        for (int i = 0; i < D1; ++i)
            O3_UNROLL
            for (int c = 0; c < DO; ++c) Pm[i][c] = 0.f;
        O3_UNROLL
        for (int uu = 0; uu < 4; ++uu)
            if (uu < nu) {
                O3_UNROLL
                for (int i = 0; i < D1; ++i) {
                    const float x = xr[uu * D1 + i];
                    O3_UNROLL
                    for (int c = 0; c < DO; ++c)
                        if ((NZ >> (i * DO + c)) & 1u) Pm[i][c] += x * Gc[uu][c];
                }
            }
        float gy[D2];
        O3_UNROLL
        for (int j = 0; j < D2; ++j) gy[j] = 0.f;
        o3_gy<L1, L2, LO>(Pm, gy);
        O3_UNROLL
        for (int j = 0; j < D2; ++j) O3_ATOMIC_ADD(gyr + j, gy[j]);
    }
}

template <int L1>
O3_DEV void o3_gin_unit(const O3Gin& S, const int32_t* GI, int ub, int e, int D1p, int D2p, int DOp) {
    constexpr int D1 = 2 * L1 + 1;
    const int32_t* tab = S.tab;
    const int u0 = 4 * ub, nu = GI[o3::GI_MUL1] - u0 < 4 ? GI[o3::GI_MUL1] - u0 : 4;
    float gx[4][D1];
    O3_UNROLL
    for (int uu = 0; uu < 4; ++uu)
        O3_UNROLL
        for (int i = 0; i < D1; ++i) gx[uu][i] = 0.f;
    const float* xr = S.xs + e * D1p + GI[o3::GI_OFF1] + u0 * D1;
    for (int r = 0; r < GI[o3::GI_NR]; ++r) {
        const int32_t* RE = tab + tab[o3::H_RE] + (GI[o3::GI_R0] + r) * o3::RE_W;
        const int32_t* IO = tab + tab[o3::H_IO] + RE[o3::RE_IO] * o3::IO_W;
        const int32_t* G = tab + tab[o3::H_GRP] + RE[o3::RE_GRP] * o3::GRP_W;
        const int mul = IO[o3::IO_MUL], KPP = 4 * IO[o3::IO_NSUB], lo = IO[o3::IO_D] >> 1;
        const float* ge = S.gs + e * DOp + IO[o3::IO_OFF];
        const float* wt0 = S.WT + tab[tab[o3::H_GIWT] + RE[o3::RE_WT0] + ub];
        for (int pi = 0; pi < G[o3::G_NP]; ++pi) {
            const int32_t* P = tab + tab[o3::H_PATH] + G[o3::G_P0 + pi] * o3::PATH_W;
            const float* yr = S.ys + e * D2p + P[o3::P_OFF2];
            float* gyr = S.need_gy ? S.gys + e * D2p + P[o3::P_OFF2] : nullptr;
            const float* wt = wt0 + 4 * pi;
            switch (P[o3::P_L2] * 3 + lo) {
#define O3_CASE(B, C)                                                                                   \
    case B * 3 + C:                                                                                     \
        if constexpr (o3_tri<L1, B, C>::v) o3_gin_path<L1, B, C>(xr, yr, gx, gyr, ge, wt, KPP, mul, nu); \
        break;
                O3_CASE(0, 0) O3_CASE(0, 1) O3_CASE(0, 2) O3_CASE(1, 0) O3_CASE(1, 1) O3_CASE(1, 2)
                O3_CASE(2, 0) O3_CASE(2, 1) O3_CASE(2, 2)
#undef O3_CASE
            }
        }
    }
    float* gxr = S.gxs + e * D1p + GI[o3::GI_OFF1] + u0 * D1;
    O3_UNROLL
    for (int uu = 0; uu < 4; ++uu)
        if (uu < nu) {
            O3_UNROLL
            for (int i = 0; i < D1; ++i) gxr[uu * D1 + i] = gx[uu][i];
        }
}

O3_DEV void o3_gin_tile(const O3Gin& S, const O3Rows& in1, const float* __restrict__ in2,
                        const float* __restrict__ gout, const O3GRows& gin1, float* __restrict__ gin2,
                        long long row0, int nrow O3_NT_DECL) {
    const int32_t* tab = S.tab;
    const int D1 = tab[o3::H_D1], D2 = tab[o3::H_D2], DO = tab[o3::H_DOUT];
    const int D1p = D1 | 1, D2p = D2 | 1, DOp = DO | 1;
    constexpr int TE = o3::TE_GIN;

    O3_THREADS
        const int warp = tid >> 5, lane = tid & 31, nw = NT >> 5;
        for (int e = warp; e < TE; e += nw) {
            const bool ok = e < nrow;
            for (int c = lane; c < D1; c += 32) S.gxs[e * D1p + c] = 0.f;   // columns without any path stay zero
            if (S.need_gy) {
                if (ok) o3_row_load(in1, row0 + e, S.xs + e * D1p, lane);
                else
                    for (int c = lane; c < D1; c += 32) S.xs[e * D1p + c] = 0.f;
            }
            if (lane < D2) {
                S.ys[e * D2p + lane] = ok ? in2[(row0 + e) * D2 + lane] : 0.f;
                S.gys[e * D2p + lane] = 0.f;
            }
            const float* gsrc = gout + (row0 + e) * DO;
            for (int c = lane; c < DO; c += 32) S.gs[e * DOp + c] = ok ? gsrc[c] : 0.f;
        }
    O3_END

    O3_THREADS
        const int warp = tid >> 5, lane = tid & 31;
        (void)NT;
        const int32_t* U = tab + tab[o3::H_GUNIT];
        for (int k = U[warp]; k < U[warp + 1]; ++k) {
            const int packed = U[o3::NWARP + 1 + k];
            const int gi = packed & 255, ub = (packed >> 8) & 0xffff, e = (packed >> 24) * 32 + lane;
            const int32_t* GI = tab + tab[o3::H_GI] + gi * o3::GI_W;
            switch (GI[o3::GI_L1]) {
                case 0: o3_gin_unit<0>(S, GI, ub, e, D1p, D2p, DOp); break;
                case 1: o3_gin_unit<1>(S, GI, ub, e, D1p, D2p, DOp); break;
                case 2: o3_gin_unit<2>(S, GI, ub, e, D1p, D2p, DOp); break;
            }
        }
    O3_END

    O3_THREADS
        o3_tile_store_grad(gin1, row0, nrow, S.gxs, D1p, tid, NT);
        const int warp = tid >> 5, lane = tid & 31, nw = NT >> 5;
        if (gin2 != nullptr)
            for (int e = warp; e < nrow; e += nw)
                if (lane < D2) gin2[(row0 + e) * D2 + lane] = S.gys[e * D2p + lane];
    O3_END
}

// (2) weight gradients.  Tile of 32 rows; per output irrep: features of ALL its channels into the shared F buffer
//     (warp = block of 4 channels, lane = row) and GT = a g transposed, then every thread adds its slice of the (e, c)
//     axis to the 4 x 4 block of weight gradients it owns FOR THE WHOLE KERNEL (register accumulators, one set per output
//     irrep, at most MAXIO_GW).  The slices are combined once, after the last tile (o3_gw_flush).
struct O3Gw {
    const int32_t* tab;
    float *xs, *ys, *gs, *F, *GT;
};

template <int L1, int L2, int LO>
O3_DEV void o3_feat_block(const float* xr, const float* yr, int nu, float* Fe, int Rp) {
    constexpr int D1 = 2 * L1 + 1, DO = 2 * LO + 1;
    constexpr unsigned NZ = o3_nz<L1, L2, LO>::mask;
    float M[D1][DO];
    o3_M<L1, L2, LO>(yr, M);
    O3_UNROLL
    for (int uu = 0; uu < 4; ++uu) {
        float fc[DO];
        O3_UNROLL
        for (int c = 0; c < DO; ++c) fc[c] = 0.f;
        if (uu < nu) {
            O3_UNROLL
            for (int i = 0; i < D1; ++i) {
                const float x = xr[uu * D1 + i];
                O3_UNROLL
                for (int c = 0; c < DO; ++c)
                    if ((NZ >> (i * DO + c)) & 1u) fc[c] += M[i][c] * x;
            }
        }
        O3_UNROLL
        for (int c = 0; c < DO; ++c) Fe[(size_t)uu * Rp + c] = fc[c];
    }
}

// features + GT of output irrep `io` (one region), then this thread's slice of its 4 x 4 block (next region)
O3_DEV void o3_gw_build(const O3Gw& S, const int32_t* IO, int tid, int NT) {
    const int32_t* tab = S.tab;
    const int D1p = tab[o3::H_D1] | 1, D2p = tab[o3::H_D2] | 1, DOp = tab[o3::H_DOUT] | 1;
    const int mul = IO[o3::IO_MUL], d = IO[o3::IO_D], nblk = IO[o3::IO_NBLK], mulp = (mul + 3) & ~3;
    const int Rp = (o3::TE_BWD * d) | 1;
    const int32_t* BL = tab + tab[o3::H_BLK] + IO[o3::IO_BLK];
    const float a = O3_I2F(IO[o3::IO_A]);
    const int warp = tid >> 5, e = tid & 31, nw = NT >> 5;
    for (int w = warp; w < mulp; w += nw)
        for (int c = 0; c < d; ++c)
            S.GT[(size_t)w * Rp + e * d + c] = w < mul ? a * S.gs[e * DOp + IO[o3::IO_OFF] + w * d + c] : 0.f;
    for (int b = warp; b < nblk; b += nw) {
        const int32_t* B = BL + b * o3::BLK_W;
        const int32_t* G = tab + tab[o3::H_GRP] + (B[o3::B_GRP] & 0xffff) * o3::GRP_W;
        const int u0 = B[o3::B_GRP] >> 16, l1 = G[o3::G_L1];
        const int nu = G[o3::G_MUL1] - u0 < 4 ? G[o3::G_MUL1] - u0 : 4;
        const float* xr = S.xs + e * D1p + G[o3::G_OFF1] + u0 * (2 * l1 + 1);
        for (int pi = 0; pi < G[o3::G_NP]; ++pi) {
            const int32_t* P = tab + tab[o3::H_PATH] + G[o3::G_P0 + pi] * o3::PATH_W;
            const float* yr = S.ys + e * D2p + P[o3::P_OFF2];
            float* Fe = S.F + (size_t)(4 * (B[o3::B_SUB0] + pi)) * Rp + e * d;
            switch (l1 * 9 + P[o3::P_L2] * 3 + (d >> 1)) {
#define O3_CASE(A, B2, C)                                    \
    case A * 9 + B2 * 3 + C:                                 \
        o3_feat_block<A, B2, C>(xr, yr, nu, Fe, Rp);         \
        break;
                O3_TRIPLES(O3_CASE)
#undef O3_CASE
            }
        }
    }
}

O3_DEV void o3_gw_accum(const O3Gw& S, const int32_t* IO, float (&acc)[16], int tid) {
    const int mul = IO[o3::IO_MUL], d = IO[o3::IO_D], nwb = ((mul + 3) & ~3) >> 2, lg = IO[o3::IO_GWLG];
    const int R = o3::TE_BWD * d, Rp = R | 1;
    const int s = tid & ((1 << lg) - 1), t = tid >> lg;
    if (t >= IO[o3::IO_NSUB] * nwb) return;
    const int sb = nwb == 1 ? t : (int)O3_MULHI((unsigned)t, (unsigned)IO[o3::IO_NWB_MAGIC]), wb = t - sb * nwb;
    const float* f = S.F + (size_t)(4 * sb) * Rp;
    const float* g = S.GT + (size_t)(4 * wb) * Rp;
    const int r0 = (s * R) >> lg, r1 = ((s + 1) * R) >> lg;
    O3_UNROLL2
    for (int r = r0; r < r1; ++r) {
        const float f0 = f[r], f1 = f[Rp + r], f2 = f[2 * Rp + r], f3 = f[3 * Rp + r];
        const float g0 = g[r], g1 = g[Rp + r], g2 = g[2 * Rp + r], g3 = g[3 * Rp + r];
        acc[0] += f0 * g0; acc[1] += f0 * g1; acc[2] += f0 * g2; acc[3] += f0 * g3;
        acc[4] += f1 * g0; acc[5] += f1 * g1; acc[6] += f1 * g2; acc[7] += f1 * g3;
        acc[8] += f2 * g0; acc[9] += f2 * g1; acc[10] += f2 * g2; acc[11] += f2 * g3;
        acc[12] += f3 * g0; acc[13] += f3 * g1; acc[14] += f3 * g2; acc[15] += f3 * g3;
    }
}

// acc[slot] belongs to the slot-th output irrep that has paths
O3_DEV void o3_gw_tile(const O3Gw& S, O3_ACC_DECL, const O3Rows& in1,
                       const float* __restrict__ in2, const float* __restrict__ gout, long long row0,
                       int nrow O3_NT_DECL) {
    const int32_t* tab = S.tab;
    const int D1 = tab[o3::H_D1], D2 = tab[o3::H_D2], DO = tab[o3::H_DOUT], nio = tab[o3::H_NIO];
    const int D1p = D1 | 1, D2p = D2 | 1, DOp = DO | 1;
    constexpr int TE = o3::TE_BWD;

    O3_THREADS
        const int warp = tid >> 5, lane = tid & 31, nw = NT >> 5;
        for (int e = warp; e < TE; e += nw) {
            const bool ok = e < nrow;
            if (ok) o3_row_load(in1, row0 + e, S.xs + e * D1p, lane);
            else
                for (int c = lane; c < D1; c += 32) S.xs[e * D1p + c] = 0.f;
            if (lane < D2) S.ys[e * D2p + lane] = ok ? in2[(row0 + e) * D2 + lane] : 0.f;
            const float* gsrc = gout + (row0 + e) * DO;
            for (int c = lane; c < DO; c += 32) S.gs[e * DOp + c] = ok ? gsrc[c] : 0.f;
        }
    O3_END

    int io = 0;
    O3_UNROLL
    for (int slot = 0; slot < o3::MAXIO_GW; ++slot) {
        while (io < nio && tab[tab[o3::H_IO] + io * o3::IO_W + o3::IO_NSUB] == 0) ++io;   // block-uniform
        if (io < nio) {
            const int32_t* IO = tab + tab[o3::H_IO] + io * o3::IO_W;
            O3_THREADS
                o3_gw_build(S, IO, tid, NT);
            O3_END
            O3_THREADS
                (void)NT;
                o3_gw_accum(S, IO, O3_ACC(acc, slot, tid), tid);
            O3_END
            ++io;
        }
    }
}

// after the last tile: combine the slices of every 4 x 4 block (through shared scratch, aliasing F) and add the block to
// the global weight gradient
O3_DEV void o3_gw_flush(const O3Gw& S, O3_ACC_DECL, float* __restrict__ gw O3_NT_DECL) {
    const int32_t* tab = S.tab;
    const int nio = tab[o3::H_NIO];
    float* scr = S.F;
    int io = 0;
    O3_UNROLL
    for (int slot = 0; slot < o3::MAXIO_GW; ++slot) {
        while (io < nio && tab[tab[o3::H_IO] + io * o3::IO_W + o3::IO_NSUB] == 0) ++io;
        if (io < nio) {
            const int32_t* IO = tab + tab[o3::H_IO] + io * o3::IO_W;
            const int mul = IO[o3::IO_MUL], nwb = ((mul + 3) & ~3) >> 2, lg = IO[o3::IO_GWLG];
            const int base = IO[o3::IO_NSUB] * nwb;
            O3_THREADS
                (void)NT;
                if ((tid >> lg) < base) {
                    O3_UNROLL
                    for (int k = 0; k < 16; ++k) scr[k * O3_SCR_LD + tid] = O3_ACC(acc, slot, tid)[k];
                }
            O3_END
            O3_THREADS
                const int32_t* BL = tab + tab[o3::H_BLK] + IO[o3::IO_BLK];
                const int32_t* SUB = tab + tab[o3::H_SUB] + IO[o3::IO_SUB];
                for (int o = tid; o < (base << 4); o += NT) {
                    const int k = o & 15, t = o >> 4;
                    float sum = 0.f;
                    for (int s = 0; s < (1 << lg); ++s) sum += scr[k * O3_SCR_LD + (t << lg) + s];
                    const int sb = nwb == 1 ? t : (int)O3_MULHI((unsigned)t, (unsigned)IO[o3::IO_NWB_MAGIC]), wb = t - sb * nwb;
                    const int word = SUB[sb];
                    const int32_t* B = BL + (word & 0xffff) * o3::BLK_W;
                    const int32_t* G = tab + tab[o3::H_GRP] + (B[o3::B_GRP] & 0xffff) * o3::GRP_W;
                    const int32_t* P = tab + tab[o3::H_PATH] + G[o3::G_P0 + (word >> 16)] * o3::PATH_W;
                    const int u = (B[o3::B_GRP] >> 16) + (k >> 2), w = 4 * wb + (k & 3);
                    if (u < G[o3::G_MUL1] && w < mul) O3_GLOBAL_ADD(gw + P[o3::P_WOFF] + u * mul + w, sum);
                }
            O3_END
            ++io;
        }
    }
}
