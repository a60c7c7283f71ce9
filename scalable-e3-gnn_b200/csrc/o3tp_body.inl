// Tile programs of the l <= 2 tensor product (forward and backward), written once and compiled twice: by nvcc inside
// o3tp.cu (one CUDA thread per `tid`) and by g++ inside tests/emu/o3tp_emu.cpp (threads run one after another, phase by
// phase) so the table walking and indexing are checked against the oracle on a machine without a GPU.
//
// The includer defines:
//   O3_DEV                     function qualifier
//   O3_THREADS / O3_END        open / close a region every thread of the block executes; O3_END is a block barrier
//   O3_ATOMIC_ADD(p, v)        shared-memory float add that tolerates several threads on one address
//   O3_I2F(i)                  reinterpret an int32 table word as float
//   O3_NT_DECL                 extra parameter `, int NT_` carrying the emulated block size (empty under nvcc)
//   o3f4 / O3_LD4(p)           four consecutive floats read from a 16-byte aligned shared-memory address
//   O3_UNROLL                  unroll pragma
//
// Math (oracle/lmax2_oracle.py forward): for every output irrep io with stacked paths,
//   F[kk, e, c]   = sum_{i,j} C_p[i,j,c] x1[e, off1_p + u d1 + i] y[e, off2_p + j]     kk = koff_p + u
//   out[e, w, c]  = a_io sum_kk F[kk, e, c] W_io[kk, w]
// backward, with g' = a_io g:
//   gW_io[kk, w]  = sum_{e,c} F[kk, e, c] g'[e, w, c]
//   G[kk, e, c]   = sum_w W_io[kk, w] g'[e, w, c]
//   gx1[e, off1 + u d1 + i] += sum C[i,j,c] y[e, j] G[koff + u, e, c];   gy[e, j] += sum C[i,j,c] x1[...] G[...]

struct O3Fwd {
    const int32_t* tab;  // table blob (shared memory)
    const float* Ws;     // all weights, per io [K, mulp] zero padded (shared, resident)
    float *xs, *ys, *os;
    int TE;
};

struct O3Bwd {
    const int32_t* tab;
    const float* WT;  // all weights transposed, per io [mul, Kp] (shared, resident)
    float* gWs;       // weight gradient accumulators, flat like the weights (shared, resident)
    float *xs, *ys, *gs, *gxs, *gys, *F, *G, *GT;
    int TE, Rp;
};

O3_DEV void o3_features(const int32_t* tab, const int32_t* IO, const float* xs, const float* ys, float* F,
                               int TE, int Rp, int D1p, int D2p, int Krows, int tid, int NT) {
    const int d = IO[o3::IO_D], K = IO[o3::IO_K];
    const int32_t* ent = tab + tab[o3::H_ENT];
    for (int item = tid; item < TE * Krows; item += NT) {
        const int kk = item / TE, e = item - kk * TE;
        float* f = F + (size_t)kk * Rp + e * d;
        if (kk >= K) {
            for (int c = 0; c < d; ++c) f[c] = 0.f;
            continue;
        }
        const int32_t* P = tab + tab[o3::H_PATH] + IO[o3::IO_PBEG] * o3::PATH_W;
        while (kk >= P[o3::P_KOFF] + P[o3::P_MUL1]) P += o3::PATH_W;
        const int u = kk - P[o3::P_KOFF];
        const float* xr = xs + e * D1p + P[o3::P_OFF1] + u * P[o3::P_D1];
        const float* yr = ys + e * D2p + P[o3::P_OFF2];
        for (int c = 0; c < d; ++c) {
            float s = 0.f;
            for (int t = P[o3::P_EB0 + c]; t < P[o3::P_EB0 + c + 1]; ++t) {
                const int32_t* E = ent + t * o3::ENT_W;
                s += O3_I2F(E[3]) * xr[E[0]] * yr[E[1]];
            }
            f[c] = s;
        }
    }
}

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

O3_DEV void o3_fwd_tile(const O3Fwd& S, const float* __restrict__ in1, const float* __restrict__ in2,
                        float* __restrict__ out, long long row0, int nrow O3_NT_DECL) {
    const int32_t* tab = S.tab;
    const int D1 = tab[o3::H_D1], D2 = tab[o3::H_D2], DO = tab[o3::H_DOUT];
    const int D1p = D1 | 1, D2p = D2 | 1, DOp = DO | 1, TE = S.TE;

    O3_THREADS
        const int warp = tid >> 5, lane = tid & 31, nw = NT >> 5;
        for (int e = warp; e < TE; e += nw) {
            const bool ok = e < nrow;
            const float* src = in1 + (row0 + e) * D1;
            for (int c = lane; c < D1; c += 32) S.xs[e * D1p + c] = ok ? src[c] : 0.f;
            if (lane < D2) S.ys[e * D2p + lane] = ok ? in2[(row0 + e) * D2 + lane] : 0.f;
        }
    O3_END

    O3_THREADS
        const int warp = tid >> 5, lane = tid & 31;
        (void)NT;
        const int32_t* U = tab + tab[o3::H_UNIT];
        for (int k = U[warp]; k < U[warp + 1]; ++k) {
            const int packed = U[o3::NWARP + 1 + k];
            const int io = packed & 255, q = (packed >> 8) & 255, e = (packed >> 16) * 32 + lane;
            const int32_t* IO = tab + tab[o3::H_IO] + io * o3::IO_W;
            const float* xe = S.xs + e * D1p;
            const float* ye = S.ys + e * D2p;
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

O3_DEV void o3_bwd_tile(const O3Bwd& S, const float* __restrict__ in1, const float* __restrict__ in2,
                               const float* __restrict__ gout, float* __restrict__ gin1, float* __restrict__ gin2,
                               long long row0, int nrow O3_NT_DECL) {
    const int32_t* tab = S.tab;
    const int D1 = tab[o3::H_D1], D2 = tab[o3::H_D2], DO = tab[o3::H_DOUT], nio = tab[o3::H_NIO];
    const int D1p = D1 | 1, D2p = D2 | 1, DOp = DO | 1, TE = S.TE, Rp = S.Rp;
    const int32_t* ent = tab + tab[o3::H_ENT];

    O3_THREADS
        for (int idx = tid; idx < TE * D1; idx += NT) {
            const int e = idx / D1, c = idx - e * D1;
            S.xs[e * D1p + c] = e < nrow ? in1[(row0 + e) * D1 + c] : 0.f;
            S.gxs[e * D1p + c] = 0.f;
        }
        for (int idx = tid; idx < TE * D2; idx += NT) {
            const int e = idx / D2, c = idx - e * D2;
            S.ys[e * D2p + c] = e < nrow ? in2[(row0 + e) * D2 + c] : 0.f;
            S.gys[e * D2p + c] = 0.f;
        }
        for (int idx = tid; idx < TE * DO; idx += NT) {
            const int e = idx / DO, c = idx - e * DO;
            S.gs[e * DOp + c] = e < nrow ? gout[(row0 + e) * DO + c] : 0.f;
        }
    O3_END

    for (int io = 0; io < nio; ++io) {
        const int32_t* IO = tab + tab[o3::H_IO] + io * o3::IO_W;
        const int mul = IO[o3::IO_MUL], d = IO[o3::IO_D], K = IO[o3::IO_K];
        const int mulp = (mul + 3) & ~3, Kp = (K + 3) & ~3, R = TE * d;
        if (K == 0) continue;  // block-uniform
        O3_THREADS
            const float a = O3_I2F(IO[o3::IO_A]);
            for (int item = tid; item < mulp * R; item += NT) {
                const int w = item / R, r = item - w * R;
                const int e = r / d, c = r - e * d;
                S.GT[(size_t)w * Rp + r] = w < mul ? a * S.gs[e * DOp + IO[o3::IO_OFF] + w * d + c] : 0.f;
            }
            o3_features(tab, IO, S.xs, S.ys, S.F, TE, Rp, D1p, D2p, Kp, tid, NT);
        O3_END
        O3_THREADS
            {   // weight gradient: 4 x 4 register blocks of F . GT^T
                const int nkb = Kp >> 2, nwb = mulp >> 2;
                for (int item = tid; item < nkb * nwb; item += NT) {
                    const int kb = item / nwb, wb = item - kb * nwb;
                    const float* f = S.F + (size_t)(4 * kb) * Rp;
                    const float* g = S.GT + (size_t)(4 * wb) * Rp;
                    float acc[4][4];
                    for (int i = 0; i < 4; ++i)
                        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
                    for (int r = 0; r < R; ++r) {
                        const float f0 = f[r], f1 = f[Rp + r], f2 = f[2 * Rp + r], f3 = f[3 * Rp + r];
                        const float g0 = g[r], g1 = g[Rp + r], g2 = g[2 * Rp + r], g3 = g[3 * Rp + r];
                        acc[0][0] += f0 * g0; acc[0][1] += f0 * g1; acc[0][2] += f0 * g2; acc[0][3] += f0 * g3;
                        acc[1][0] += f1 * g0; acc[1][1] += f1 * g1; acc[1][2] += f1 * g2; acc[1][3] += f1 * g3;
                        acc[2][0] += f2 * g0; acc[2][1] += f2 * g1; acc[2][2] += f2 * g2; acc[2][3] += f2 * g3;
                        acc[3][0] += f3 * g0; acc[3][1] += f3 * g1; acc[3][2] += f3 * g2; acc[3][3] += f3 * g3;
                    }
                    float* gw = S.gWs + IO[o3::IO_WOFF];
                    O3_UNROLL
                    for (int i = 0; i < 4; ++i)
                        O3_UNROLL
                        for (int j = 0; j < 4; ++j) {
                            const int kk = 4 * kb + i, w = 4 * wb + j;
                            if (kk < K && w < mul) gw[kk * mul + w] += acc[i][j];
                        }
                }
            }
            {   // G = W . g'
                const int nkc = Kp >> 2;
                for (int item = tid; item < R * nkc; item += NT) {
                    const int kc = item / R, r = item - kc * R;
                    const float* wt = S.WT + IO[o3::IO_WTOFF] + 4 * kc;
                    const float* g = S.GT + r;
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                    for (int w = 0; w < mul; ++w) {
                        const float gv = g[(size_t)w * Rp];
                        const o3f4 wv = O3_LD4(wt + w * Kp);
                        a0 += gv * wv.x; a1 += gv * wv.y; a2 += gv * wv.z; a3 += gv * wv.w;
                    }
                    float* G = S.G + (size_t)(4 * kc) * Rp + r;
                    G[0] = a0; G[Rp] = a1; G[2 * Rp] = a2; G[3 * Rp] = a3;
                }
            }
        O3_END
        O3_THREADS
            for (int gi = IO[o3::IO_GBEG]; gi < IO[o3::IO_GEND]; ++gi) {
                const int32_t* Gp = tab + tab[o3::H_GRP] + gi * o3::GRP_W;
                const int off1 = Gp[o3::G_OFF1], d1 = Gp[o3::G_D1], mul1 = Gp[o3::G_MUL1], nj = Gp[o3::G_NJ];
                for (int item = tid; item < TE * mul1; item += NT) {
                    const int u = item / TE, e = item - u * TE;
                    const float* yr = S.ys + e * D2p;
                    const float* xr = S.xs + e * D1p + off1 + u * d1;
                    float* gx = S.gxs + e * D1p + off1 + u * d1;
                    const float* Gr = S.G + (size_t)u * Rp + e * d;
                    for (int i = 0; i < d1; ++i) {
                        float s = 0.f;
                        for (int t = Gp[o3::G_IB0 + i]; t < Gp[o3::G_IB0 + i + 1]; ++t) {
                            const int32_t* E = ent + t * o3::ENT_W;
                            s += O3_I2F(E[3]) * yr[E[0]] * Gr[(size_t)E[1] * Rp + E[2]];
                        }
                        gx[i] += s;
                    }
                    if (gin2 != nullptr) {
                        for (int jj = 0; jj < nj; ++jj) {
                            float s = 0.f;
                            for (int t = Gp[o3::G_JB0 + jj]; t < Gp[o3::G_JB0 + jj + 1]; ++t) {
                                const int32_t* E = ent + t * o3::ENT_W;
                                s += O3_I2F(E[3]) * xr[E[0]] * Gr[(size_t)E[1] * Rp + E[2]];
                            }
                            O3_ATOMIC_ADD(&S.gys[e * D2p + Gp[o3::G_JABS + jj]], s);
                        }
                    }
                }
            }
        O3_END
    }

    O3_THREADS
        for (int idx = tid; idx < nrow * D1; idx += NT) {
            const int e = idx / D1, c = idx - e * D1;
            gin1[(row0 + e) * D1 + c] = S.gxs[e * D1p + c];
        }
        if (gin2 != nullptr)
            for (int idx = tid; idx < nrow * D2; idx += NT) {
                const int e = idx / D2, c = idx - e * D2;
                gin2[(row0 + e) * D2 + c] = S.gys[e * D2p + c];
            }
    O3_END
}
