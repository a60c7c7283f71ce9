// TEST INFRASTRUCTURE ONLY.  CPU emulation of the o3tp CUDA tile programs: includes the very source the kernels are
// built from (csrc/o3tp_body.inl + csrc/o3tp_tables.h) with the block's threads run one after another, phase by phase,
// and mirrors the kernel wrappers' shared-memory carve-up and weight staging (csrc/o3tp.cu).  Lets the CPU test suite
// check the table walking / indexing of the GPU code against the oracle; it is never used by the product path.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../scalable-e3-gnn_b200/csrc/o3tp_tables.h"

namespace {
struct o3f4 { float x, y, z, w; };
inline float i2f(int32_t i) { float f; std::memcpy(&f, &i, 4); return f; }
#define O3_DEV inline
// thread order inside a region: a region must not depend on it (no barrier-free read-after-write between threads), so the
// tests run every case in several orders
static int g_order = 0;
static inline int emu_tid(int k, int nt) {
    if (g_order == 1) return nt - 1 - k;                                  // reversed
    if (g_order == 2) return ((k & 31) * (nt >> 5) + (k >> 5)) % nt;      // lane-major: warps interleaved
    if (g_order == 3) return (k * 37 + 11) % nt;                          // a fixed permutation (37 coprime to 256)
    return k;
}
#define O3_THREADS for (int k_ = 0; k_ < NT_; ++k_) { const int tid = emu_tid(k_, NT_); const int NT = NT_;
#define O3_END }
#define O3_ATOMIC_ADD(p, v) (*(p) += (v))
#define O3_GW_ADD(S, p, v) (*(p) += (v))
#define O3_MULHI(a, b) ((unsigned)(((unsigned long long)(a) * (unsigned long long)(b)) >> 32))
#define O3_I2F(i) i2f(i)
#define O3_NT_DECL , int NT_
#define O3_LD4(p) (o3f4{(p)[0], (p)[1], (p)[2], (p)[3]})
#define O3_UNROLL
#define O3_UNROLL2
#define O3_ACC_DECL float (*acc)[o3::MAXIO_GW][16]
#define O3_ACC(acc, slot, tid) acc[tid][slot]
#define O3_GLOBAL_ADD(p, v) (*(p) += (v))
#define O3_GLOBAL_ADD4(p, a, b, c, d) ((p)[0] += (a), (p)[1] += (b), (p)[2] += (c), (p)[3] += (d))
#define O3_CP4(dst, src) (*(dst) = *(src))
#define O3_CP_COMMIT()
#define O3_CP_WAIT()
#define __restrict__
#include "../../scalable-e3-gnn_b200/csrc/o3tp_cg_gen.inl"
#include "../../scalable-e3-gnn_b200/csrc/o3tp_body.inl"
}  // namespace

static void make_rows(int nseg, const float* const* base, const int* const* idx, const int* width, const int* ld,
                      float* const* gbase, const int* gmode, O3Rows& X, O3GRows& G) {
    int c0 = 0;
    X.nseg = G.nseg = nseg;
    for (int s = 0; s < 4; ++s) {
        const bool on = s < nseg;
        X.base[s] = on ? base[s] : nullptr; X.idx[s] = on ? idx[s] : nullptr;
        X.ld[s] = on ? ld[s] : 0; X.c0[s] = c0; X.width[s] = on ? width[s] : 0;
        G.base[s] = on && gbase ? gbase[s] : nullptr; G.idx[s] = X.idx[s]; G.ld[s] = X.ld[s]; G.c0[s] = c0;
        G.width[s] = X.width[s]; G.mode[s] = on && gbase && gbase[s] ? gmode[s] : 0;
        c0 += X.width[s];
    }
}

static bool make_plan(o3::Plan& P, int n1, const int* in1, int n2, const int* in2, int no, const int* out) {
    for (int i = 0; i < n1; ++i) P.in1.push_back({in1[3 * i], in1[3 * i + 1], in1[3 * i + 2]});
    for (int i = 0; i < n2; ++i) P.in2.push_back({1, in2[2 * i], in2[2 * i + 1]});
    for (int i = 0; i < no; ++i) P.out.push_back({out[3 * i], out[3 * i + 1], out[3 * i + 2]});
    return o3::build_plan(P);
}

extern "C" {

void emu_set_order(int mode) { g_order = mode; }

// irreps as flat int triples (mul, l, p) / pairs (l, p); returns weight count or -1; dims = D1, D2, Dout, npaths
int emu_plan(int n1, const int* in1, int n2, const int* in2, int no, const int* out, int* dims, int* path_io,
             int* path_i1, int* path_i2, int* path_woff, float* path_a) {
    o3::Plan P;
    if (!make_plan(P, n1, in1, n2, in2, no, out)) return -1;
    dims[0] = P.D1; dims[1] = P.D2; dims[2] = P.Dout; dims[3] = (int)P.paths.size();
    for (size_t k = 0; k < P.paths.size(); ++k) {
        path_io[k] = P.paths[k].io; path_i1[k] = P.paths[k].i1; path_i2[k] = P.paths[k].i2;
        path_woff[k] = P.paths[k].woff; path_a[k] = P.a[P.paths[k].io];
    }
    return P.nW;
}

int emu_coupling(int l1, int l2, int l3, double* out) {
    double C[5][5][5];
    if (!o3::cg(l1, l2, l3, C)) return -1;
    for (int i = 0; i < 2 * l1 + 1; ++i)
        for (int j = 0; j < 2 * l2 + 1; ++j)
            for (int k = 0; k < 2 * l3 + 1; ++k) *out++ = C[i][j][k];
    return 0;
}

int emu_forward(int n1, const int* in1i, int n2, const int* in2i, int no, const int* outi, long long rows, int nseg,
                const float* const* sbase, const int* const* sidx, const int* swidth, const int* sld, const float* in2,
                const float* w, float* out, int TE, int NT, int nblocks) {
    O3Rows in1; O3GRows unusedG;
    make_rows(nseg, sbase, sidx, swidth, sld, nullptr, nullptr, in1, unusedG);
    o3::Plan P;
    if (!make_plan(P, n1, in1i, n2, in2i, no, outi)) return -1;
    if (TE % 32 || NT != 32 * o3::NWARP) return -2;  // the forward schedule is made for 8 warps and 32-row groups
    o3::schedule_forward(P, TE);
    const int32_t* tab = P.blob.data();
    const long long ntiles = (rows + TE - 1) / TE;
    for (int b = 0; b < nblocks; ++b) {
        std::vector<float> sm(o3::fwd_floats(P.blob, TE), -1e30f);  // poison: unwritten reads show up
        float* fl = sm.data();
        float* Ws = fl; fl += tab[o3::H_NWP];
        O3Fwd S;
        S.tab = tab; S.Ws = Ws; S.TE = TE;
        S.xs0 = fl; fl += TE * (tab[o3::H_D1] | 1);
        S.xs1 = fl; fl += TE * (tab[o3::H_D1] | 1);
        S.ys0 = fl; fl += TE * (tab[o3::H_D2] | 1);
        S.ys1 = fl; fl += TE * (tab[o3::H_D2] | 1);
        S.os = fl;
        if (b < ntiles)
            for (int tid = 0; tid < NT; ++tid)
                o3_fwd_load(S, 0, in1, in2, (long long)b * TE, (int)std::min<long long>(TE, rows - (long long)b * TE), tid, NT);
        for (int io = 0; io < tab[o3::H_NIO]; ++io) {
            const int32_t* IO = tab + tab[o3::H_IO] + io * o3::IO_W;
            const int mul = IO[o3::IO_MUL], K = IO[o3::IO_K], mulp = IO[o3::IO_MULP];
            for (int idx = 0; idx < K * mulp; ++idx) {
                const int kk = idx / mulp, c = idx - kk * mulp;
                Ws[IO[o3::IO_WSOFF] + idx] = c < mul ? w[IO[o3::IO_WOFF] + kk * mul + c] : 0.f;
            }
        }
        int buf = 0;
        for (long long tile = b; tile < ntiles; tile += nblocks, buf ^= 1) {
            const long long row0 = tile * TE, next = tile + nblocks;
            const int nrow_next = next < ntiles ? (int)std::min<long long>(TE, rows - next * TE) : 0;
            o3_fwd_tile(S, buf, in1, in2, out, row0, (int)std::min<long long>(TE, rows - row0), next * TE, nrow_next, NT);
        }
    }
    return 0;
}

int emu_backward(int n1, const int* in1i, int n2, const int* in2i, int no, const int* outi, long long rows, int nseg,
                 const float* const* sbase, const int* const* sidx, const int* swidth, const int* sld, float* const* gbase,
                 const int* gmode, const float* in2, const float* w, const float* gout, float* gin2, float* gw, int NT,
                 int nblocks) {
    O3Rows in1; O3GRows gin1;
    make_rows(nseg, sbase, sidx, swidth, sld, gbase, gmode, in1, gin1);
    o3::Plan P;
    if (!make_plan(P, n1, in1i, n2, in2i, no, outi)) return -1;
    if (NT != 32 * o3::NWARP) return -2;  // one block of <= 4 channels per warp and round
    const int32_t* tab = P.blob.data();
    constexpr int TE = o3::TE_BWD;
    const long long ntiles = (rows + TE - 1) / TE;
    for (int i = 0; i < P.nW; ++i) gw[i] = 0.f;
    for (int b = 0; b < nblocks; ++b) {
        std::vector<float> sm(o3::bwd_floats(P.blob), -1e30f);
        float* fl = sm.data();
        float* WT = fl; fl += tab[o3::H_NWT];
        float* scr = fl; fl += 16 * O3_SCR_LD;
        float* gWs = fl; fl += tab[o3::H_NW];
        const int D1p = tab[o3::H_D1] | 1, D2p = tab[o3::H_D2] | 1, DOp = tab[o3::H_DOUT] | 1;
        O3Bwd S;
        S.tab = tab; S.WT = WT; S.gWs = gWs; S.gw_global = 0;
        S.xs0 = fl; fl += TE * D1p;
        S.xs1 = fl; fl += TE * D1p;
        S.gxs = fl; fl += TE * D1p;
        S.ys0 = fl; fl += TE * D2p;
        S.ys1 = fl; fl += TE * D2p;
        S.gys = fl; fl += TE * D2p;
        S.gs0 = fl; fl += TE * DOp;
        S.gs1 = fl; fl += TE * DOp;
        S.F = fl; fl += (size_t)4 * tab[o3::H_MAXNP] * o3::NWARP * tab[o3::H_FROW];
        S.GT = fl;
        S.scr = scr;
        for (int io = 0; io < tab[o3::H_NIO]; ++io) {
            const int32_t* IO = tab + tab[o3::H_IO] + io * o3::IO_W;
            const int32_t* BL = tab + tab[o3::H_BLK] + IO[o3::IO_BLK];
            const int32_t* SUB = tab + tab[o3::H_SUB] + IO[o3::IO_SUB];
            const int mul = IO[o3::IO_MUL], KPP = 4 * IO[o3::IO_NSUB];
            for (int idx = 0; idx < mul * KPP; ++idx) {
                const int wi = idx / KPP, kkp = idx - wi * KPP, word = SUB[kkp >> 2];
                const int32_t* B = BL + (word & 0xffff) * o3::BLK_W;
                const int32_t* G = tab + tab[o3::H_GRP] + (B[o3::B_GRP] & 0xffff) * o3::GRP_W;
                const int32_t* PP = tab + tab[o3::H_PATH] + G[o3::G_P0 + (word >> 16)] * o3::PATH_W;
                const int u = (B[o3::B_GRP] >> 16) + (kkp & 3);
                WT[IO[o3::IO_WTOFF] + idx] = u < G[o3::G_MUL1] ? w[PP[o3::P_WOFF] + u * mul + wi] : 0.f;
            }
        }
        for (int idx = 0; idx < tab[o3::H_NW]; ++idx) gWs[idx] = 0.f;
        if (b < ntiles)
            for (int tid = 0; tid < NT; ++tid)
                o3_bwd_load(S, 0, in1, in2, gout, (long long)b * TE, (int)std::min<long long>(TE, rows - (long long)b * TE),
                            tid, NT);
        int buf = 0;
        for (long long tile = b; tile < ntiles; tile += nblocks, buf ^= 1) {
            const long long row0 = tile * TE, next = tile + nblocks;
            const int nrow_next = next < ntiles ? (int)std::min<long long>(TE, rows - next * TE) : 0;
            o3_bwd_tile(S, buf, in1, in2, gout, gin1, gin2, row0, (int)std::min<long long>(TE, rows - row0), next * TE,
                        nrow_next, NT);
        }
        for (int idx = 0; idx < tab[o3::H_NW]; ++idx) gw[idx] += gWs[idx];
    }
    return 0;
}

// the split backward: input-gradient kernel, then weight-gradient kernel (mirrors o3tp_gin_kernel / o3tp_gw_kernel);
// returns 1 if the plan does not allow the split (too many output irreps / blocks)
int emu_backward_split(int n1, const int* in1i, int n2, const int* in2i, int no, const int* outi, long long rows, int nseg,
                       const float* const* sbase, const int* const* sidx, const int* swidth, const int* sld,
                       float* const* gbase, const int* gmode, const float* in2, const float* w, const float* gout,
                       float* gin2, float* gw, int NT, int nblocks) {
    O3Rows in1; O3GRows gin1;
    make_rows(nseg, sbase, sidx, swidth, sld, gbase, gmode, in1, gin1);
    o3::Plan P;
    if (!make_plan(P, n1, in1i, n2, in2i, no, outi)) return -1;
    if (NT != 32 * o3::NWARP) return -2;
    o3::schedule_forward(P, 64);
    const int32_t* tab = P.blob.data();
    if (!tab[o3::H_SPLIT]) return 1;
    for (int i = 0; i < P.nW; ++i) gw[i] = 0.f;
    {
        constexpr int TE = o3::TE_GIN;
        const long long ntiles = (rows + TE - 1) / TE;
        for (int b = 0; b < nblocks; ++b) {
            std::vector<float> sm(o3::gin_floats(P.blob), -1e30f);
            float* fl = sm.data();
            float* WT = fl; fl += tab[o3::H_NWT];
            const int D1p = tab[o3::H_D1] | 1, D2p = tab[o3::H_D2] | 1;
            O3Gin S;
            S.tab = tab; S.WT = WT; S.need_gy = gin2 != nullptr;
            S.xs = fl; fl += TE * D1p;
            S.gxs = fl; fl += TE * D1p;
            S.ys = fl; fl += TE * D2p;
            S.gys = fl; fl += TE * D2p;
            S.gs = fl;
            for (int io = 0; io < tab[o3::H_NIO]; ++io) {
                const int32_t* IO = tab + tab[o3::H_IO] + io * o3::IO_W;
                const int32_t* BL = tab + tab[o3::H_BLK] + IO[o3::IO_BLK];
                const int32_t* SUB = tab + tab[o3::H_SUB] + IO[o3::IO_SUB];
                const int mul = IO[o3::IO_MUL], KPP = 4 * IO[o3::IO_NSUB];
                const float a = i2f(IO[o3::IO_A]);
                for (int idx = 0; idx < mul * KPP; ++idx) {
                    const int wi = idx / KPP, kkp = idx - wi * KPP, word = SUB[kkp >> 2];
                    const int32_t* B = BL + (word & 0xffff) * o3::BLK_W;
                    const int32_t* G = tab + tab[o3::H_GRP] + (B[o3::B_GRP] & 0xffff) * o3::GRP_W;
                    const int32_t* PP = tab + tab[o3::H_PATH] + G[o3::G_P0 + (word >> 16)] * o3::PATH_W;
                    const int u = (B[o3::B_GRP] >> 16) + (kkp & 3);
                    WT[IO[o3::IO_WTOFF] + idx] = u < G[o3::G_MUL1] ? a * w[PP[o3::P_WOFF] + u * mul + wi] : 0.f;
                }
            }
            for (long long tile = b; tile < ntiles; tile += nblocks) {
                const long long row0 = tile * TE;
                o3_gin_tile(S, in1, in2, gout, gin1, gin2, row0, (int)std::min<long long>(TE, rows - row0), NT);
            }
        }
    }
    {
        constexpr int TE = o3::TE_BWD;
        const long long ntiles = (rows + TE - 1) / TE;
        for (int b = 0; b < nblocks; ++b) {
            std::vector<float> sm(o3::gw_floats(P.blob), -1e30f);
            float* fl = sm.data();
            O3Gw S;
            S.tab = tab;
            S.F = fl; fl += std::max(tab[o3::H_FMAX], 16 * O3_SCR_LD);
            S.GT = fl; fl += tab[o3::H_GTMAX];
            S.xs = fl; fl += TE * (tab[o3::H_D1] | 1);
            S.ys = fl; fl += TE * (tab[o3::H_D2] | 1);
            S.gs = fl;
            std::vector<float> accs((size_t)NT * o3::MAXIO_GW * 16, 0.f);
            auto* acc = reinterpret_cast<float (*)[o3::MAXIO_GW][16]>(accs.data());
            for (long long tile = b; tile < ntiles; tile += nblocks) {
                const long long row0 = tile * TE;
                o3_gw_tile(S, acc, in1, in2, gout, row0, (int)std::min<long long>(TE, rows - row0), NT);
            }
            o3_gw_flush(S, acc, gw, NT);
        }
    }
    return 0;
}

}  // extern "C"
