// Host-side planning of the l <= 2 tensor product (BASELINE configs[2]): coupling tensors, path list,
// normalisation and the flat int32 table ("blob") the tile programs in o3tp_body.inl walk.  Pure C++ (no CUDA) so the
// CPU emulation of the tile programs (tests/emu) shares it with the library.
//
// Specification: oracle/lmax2_oracle.py.  For l <= 1 the couplings and the normalisation are the reference's
// (/root/reference/models/segnn/l1_tensor_prod.py:91-94 constants, :122-189 'component' x 'element'); the reference has
// no l = 2 code (L1TP:13-14), so l = 2 follows the same convention: unit-norm invariant tensors, real bases
// l=1 -> (x,y,z), l=2 -> Q_a below.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace o3 {

enum { MAX_IRR = 8 };
// blob header words
enum { H_NIO = 0, H_D1, H_D2, H_DOUT, H_IO, H_PATH, H_BLK, H_WORDS, H_NW, H_NPATH, H_NWP, H_NWT, H_UNIT, H_TEF, H_BASE,
       H_FROW, H_GTMAX, H_GRP, H_SUB, H_MAXNP, H_GI, H_NGI, H_RE, H_GIWT, H_GUNIT, H_SPLIT, H_FMAX, HDR_W = 32 };
enum { TE_GIN = 64 };  // rows per tile of the input-gradient kernel (two 32-row groups)
enum { NWARP = 8 };   // warps per CTA both schedules are made for
enum { TE_BWD = 32 }; // rows per backward tile (lane = row)
// per output irrep
// IO_CW: output channels per forward work unit (4, 8 or 12); IO_MULP: mul padded to a multiple of IO_CW (row length of
// the staged forward weights at IO_WSOFF); IO_NQ = IO_MULP / IO_CW.
// Backward: the paths into an output irrep are grouped by input irrep (GRP record: <= 3 paths, one per in2 irrep).  A
// block = <= 4 channels of one group (all its paths): BLK record = (group | first channel << 16, first sub-block).  A
// sub-block = 4 feature rows of one (block, path): word [H_SUB + IO_SUB + s] = block | path slot << 16.  The transposed
// weights at IO_WTOFF are [mul, 4 * IO_NSUB] in sub-block order.
enum { IO_MUL = 0, IO_D, IO_OFF, IO_K, IO_WOFF, IO_A, IO_PBEG, IO_PEND, IO_WSOFF, IO_WTOFF, IO_CW, IO_MULP, IO_NQ,
       IO_NBLK, IO_BLK, IO_NSUB, IO_SUB, IO_NWB_MAGIC /* ceil(2^32 / ceil(mul / 4)) */,
       IO_GWLG /* log2 of the (e, c) slices per 4 x 4 block in the weight-gradient kernel */, IO_W = 20 };
enum { P_OFF1 = 0, P_D1, P_MUL1, P_OFF2, P_KOFF, P_L1, P_L2, P_WOFF, PATH_W };
enum { G_OFF1 = 0, G_L1, G_MUL1, G_NP, G_P0, GRP_W = 8 };
enum { B_GRP = 0, B_SUB0, BLK_W };
enum { MAXP = 3 };  // paths per group = in2 irreps
// split backward, input-gradient kernel: GI record per in1 irrep that has paths, its reach entries (output irrep, group,
// index of the per-channel-block offsets of the transposed weights)
enum { GI_OFF1 = 0, GI_L1, GI_MUL1, GI_NR, GI_R0, GI_W = 6 };
enum { RE_IO = 0, RE_GRP, RE_WT0, RE_W = 4 };
enum { MAXIO_GW = 4 };  // output irreps the weight-gradient kernel keeps register accumulators for

struct Irrep { int mul, l, p; };
struct PathH { int i1, i2, io, woff, koff; };

inline double eps3(int i, int j, int k) {
    if (i == j || j == k || i == k) return 0.0;
    return ((j - i + 3) % 3 == 1) ? 1.0 : -1.0;
}

// orthonormal (Frobenius) 3x3 bases: l=0 identity, l=1 antisymmetric, l=2 symmetric traceless Q_a in the order
// (xy, yz, 2zz-xx-yy, zx, xx-yy)
inline void basis(int l, double B[5][3][3]) {
    std::memset(B, 0, sizeof(double) * 45);
    const double s2 = 1.0 / std::sqrt(2.0), s3 = 1.0 / std::sqrt(3.0), s6 = 1.0 / std::sqrt(6.0);
    if (l == 0) {
        for (int a = 0; a < 3; ++a) B[0][a][a] = s3;
    } else if (l == 1) {
        for (int i = 0; i < 3; ++i)
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) B[i][a][b] = eps3(i, a, b) * s2;
    } else {
        B[0][0][1] = B[0][1][0] = s2;
        B[1][1][2] = B[1][2][1] = s2;
        B[2][0][0] = B[2][1][1] = -s6; B[2][2][2] = 2 * s6;
        B[3][0][2] = B[3][2][0] = s2;
        B[4][0][0] = s2; B[4][1][1] = -s2;
    }
}

// Unit-norm coupling tensor C[i][j][k] of l1 x l2 -> l3.  With a scalar factor it is +delta/sqrt(dim) (the reference's
// cg000 / cg110 / cg011); otherwise trace(B1_i B2_j B3_k), which for 1x1->1 is +epsilon/sqrt6 (the reference's cg111 with
// out = in1 x in2, L1TP:279).  Returns false if the triangle rule excludes the triple.
inline bool cg(int l1, int l2, int l3, double C[5][5][5]) {
    std::memset(C, 0, sizeof(double) * 125);
    if (l1 < 0 || l2 < 0 || l3 < 0 || l1 > 2 || l2 > 2 || l3 > 2) return false;
    if (l3 < std::abs(l1 - l2) || l3 > l1 + l2) return false;
    if (l1 == 0) {
        for (int k = 0; k < 2 * l3 + 1; ++k) C[0][k][k] = 1.0 / std::sqrt(2.0 * l3 + 1);
        return true;
    }
    if (l2 == 0) {
        for (int k = 0; k < 2 * l3 + 1; ++k) C[k][0][k] = 1.0 / std::sqrt(2.0 * l3 + 1);
        return true;
    }
    if (l3 == 0) {
        for (int k = 0; k < 2 * l1 + 1; ++k) C[k][k][0] = 1.0 / std::sqrt(2.0 * l1 + 1);
        return true;
    }
    double B1[5][3][3], B2[5][3][3], B3[5][3][3];
    basis(l1, B1); basis(l2, B2); basis(l3, B3);
    double nrm = 0;
    for (int i = 0; i < 2 * l1 + 1; ++i)
        for (int j = 0; j < 2 * l2 + 1; ++j)
            for (int k = 0; k < 2 * l3 + 1; ++k) {
                double s = 0;
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b)
                        for (int c = 0; c < 3; ++c) s += B1[i][a][b] * B2[j][b][c] * B3[k][c][a];
                if (std::fabs(s) < 1e-14) s = 0;
                C[i][j][k] = s;
                nrm += s * s;
            }
    nrm = std::sqrt(nrm);
    for (int i = 0; i < 5; ++i)
        for (int j = 0; j < 5; ++j)
            for (int k = 0; k < 5; ++k) C[i][j][k] /= nrm;
    return true;
}

struct Plan {
    std::vector<Irrep> in1, in2, out;
    std::vector<PathH> paths;
    std::vector<float> a;  // per output irrep
    std::vector<int32_t> blob;
    int D1 = 0, D2 = 0, Dout = 0, nW = 0;
    std::string err;
};

inline int32_t f2i(float f) { int32_t i; std::memcpy(&i, &f, 4); return i; }

// paths in the reference's enumeration order (io, i2, i1) (L1TP:122-151), triangle + parity selection rules; weights of
// path p are a [mul1, mul_out] row-major block at woff (blocks of one output irrep are adjacent: stacked [K, mul_out]).
inline bool build_plan(Plan& P) {
    if (P.in1.empty() || P.out.empty() || P.in2.empty() || (int)P.in1.size() > MAX_IRR || (int)P.out.size() > MAX_IRR ||
        (int)P.in2.size() > 3) {
        P.err = "o3tp: 1..8 in1/out irreps and 1..3 in2 irreps supported";
        return false;
    }
    std::vector<int> off1, off2, offo;
    auto offs = [&](const std::vector<Irrep>& v, std::vector<int>& o, bool mul1) {
        int acc = 0;
        for (auto& ir : v) {
            if (ir.l < 0 || ir.l > 2 || (ir.p != 1 && ir.p != -1) || ir.mul < 1 || (mul1 && ir.mul != 1)) return -1;
            o.push_back(acc);
            acc += ir.mul * (2 * ir.l + 1);
        }
        return acc;
    };
    P.D1 = offs(P.in1, off1, false);
    P.D2 = offs(P.in2, off2, true);
    P.Dout = offs(P.out, offo, false);
    if (P.D1 < 0 || P.D2 < 0 || P.Dout < 0) {
        P.err = "o3tp: irreps need 0 <= l <= 2, parity +-1, mul >= 1 (in2: mul == 1)";
        return false;
    }
    P.paths.clear();
    P.a.assign(P.out.size(), 0.f);
    std::vector<int32_t> io_w, path_w, blk_w, grp_w, sub_w;
    struct GrpH { int io, i1, g, blk0, nblk, wtoff; };
    std::vector<GrpH> grps;
    bool split_ok = true;
    int nio_active = 0, fmax = 4;
    int woff = 0, wsoff = 0, wtoff = 0, frow = 1, gtmax = 4, maxnp = 1;
    for (size_t io = 0; io < P.out.size(); ++io) {
        const Irrep o = P.out[io];
        const int d = 2 * o.l + 1;
        const int pbeg = (int)P.paths.size();
        int K = 0;
        const int woff0 = woff;
        for (size_t i2 = 0; i2 < P.in2.size(); ++i2)
            for (size_t i1 = 0; i1 < P.in1.size(); ++i1) {
                const Irrep a = P.in1[i1], b = P.in2[i2];
                if (o.l < std::abs(a.l - b.l) || o.l > a.l + b.l || o.p != a.p * b.p) continue;
                P.paths.push_back({(int)i1, (int)i2, (int)io, woff, K});
                woff += a.mul * o.mul;
                K += a.mul;
            }
        const int pend = (int)P.paths.size();
        P.a[io] = K > 0 ? (float)std::sqrt((double)d / (double)K) : 0.f;
        const int blk0 = (int)(blk_w.size() / BLK_W), sub0 = (int)sub_w.size();
        struct Blk { int g, u0, np; double cost; };
        std::vector<Blk> pending;
        for (int p = pbeg; p < pend; ++p) {
            const PathH& ph = P.paths[p];
            const Irrep a = P.in1[ph.i1], b = P.in2[ph.i2];
            int32_t rec[PATH_W] = {0};
            rec[P_OFF1] = off1[ph.i1]; rec[P_D1] = 2 * a.l + 1; rec[P_MUL1] = a.mul; rec[P_OFF2] = off2[ph.i2];
            rec[P_KOFF] = ph.koff; rec[P_L1] = a.l; rec[P_L2] = b.l; rec[P_WOFF] = ph.woff;
            path_w.insert(path_w.end(), rec, rec + PATH_W);
        }
        for (size_t i1 = 0; i1 < P.in1.size(); ++i1) {
            int32_t rec[GRP_W] = {0};
            int np = 0;
            for (int p = pbeg; p < pend; ++p)
                if (P.paths[p].i1 == (int)i1) rec[G_P0 + np++] = p;
            if (np == 0) continue;
            maxnp = std::max(maxnp, np);
            rec[G_OFF1] = off1[i1]; rec[G_L1] = P.in1[i1].l; rec[G_MUL1] = P.in1[i1].mul; rec[G_NP] = np;
            const int g = (int)(grp_w.size() / GRP_W);
            grp_w.insert(grp_w.end(), rec, rec + GRP_W);
            grps.push_back({(int)io, (int)i1, g, blk0, 0, wtoff});
            // instruction estimate of one block: per path the G contraction over mul outputs + 4 channels of coupling work
            const int d1 = 2 * P.in1[i1].l + 1;
            const double cost = np * (o.mul * (5.0 * d + 2) + 4.0 * (d1 + 3.0 * d1 * d + d) + 40);
            for (int u0 = 0; u0 < P.in1[i1].mul; u0 += 4) pending.push_back({g, u0, np, cost});
        }
        // heaviest blocks first, so that the NWARP blocks of a round cost about the same (a round ends at a barrier)
        std::stable_sort(pending.begin(), pending.end(), [](const Blk& x, const Blk& y) { return x.cost > y.cost; });
        for (const Blk& k : pending) {
            const int b = (int)(blk_w.size() / BLK_W) - blk0;
            blk_w.push_back(k.g | (k.u0 << 16));
            blk_w.push_back((int)sub_w.size() - sub0);
            for (int pi = 0; pi < k.np; ++pi) sub_w.push_back(b | (pi << 16));
        }
        const int nblk = (int)(blk_w.size() / BLK_W) - blk0, nsub = (int)sub_w.size() - sub0;
        // forward chunk width: least padding among {12, 8, 4} (ties: the wider), accumulators d * CW <= 40 registers
        int cw = 4, best = 1 << 30;
        for (int c : {12, 8, 4}) {
            if (c * d > 40) continue;
            const int padded = (o.mul + c - 1) / c * c;
            if (padded < best) { best = padded; cw = c; }
        }
        int32_t rec[IO_W] = {0};
        rec[IO_CW] = cw; rec[IO_MULP] = best; rec[IO_NQ] = best / cw;
        rec[IO_MUL] = o.mul; rec[IO_D] = d; rec[IO_OFF] = offo[io]; rec[IO_K] = K; rec[IO_WOFF] = woff0;
        rec[IO_A] = f2i(P.a[io]); rec[IO_PBEG] = pbeg; rec[IO_PEND] = pend; rec[IO_WSOFF] = wsoff;
        rec[IO_WTOFF] = wtoff; rec[IO_NBLK] = nblk; rec[IO_BLK] = blk0 * BLK_W; rec[IO_NSUB] = nsub; rec[IO_SUB] = sub0;
        {   // weight-gradient kernel: one 4 x 4 block and one slice per thread
            const int base = nsub * ((o.mul + 3) / 4);
            int lg = 0;
            while (lg < 5 && (base << (lg + 1)) <= 32 * NWARP) ++lg;
            rec[IO_GWLG] = lg;
            if (base > 32 * NWARP) split_ok = false;
            if (nsub > 0) ++nio_active;
            fmax = std::max(fmax, 4 * nsub * ((TE_BWD * d) | 1));
        }
        rec[IO_NWB_MAGIC] = (int32_t)(uint32_t)((0x100000000ull + (uint64_t)((o.mul + 3) / 4) - 1) / (uint64_t)((o.mul + 3) / 4));
        io_w.insert(io_w.end(), rec, rec + IO_W);
        wsoff += K * best;
        wtoff += o.mul * 4 * nsub;
        const int rp = (TE_BWD * d) | 1;
        frow = std::max(frow, rp);
        gtmax = std::max(gtmax, ((o.mul + 3) & ~3) * rp);
    }
    if (nio_active > MAXIO_GW) split_ok = false;
    // input-gradient kernel tables
    std::vector<int32_t> gi_w, re_w, giwt_w;
    for (size_t i1 = 0; i1 < P.in1.size(); ++i1) {
        int32_t rec[GI_W] = {0};
        rec[GI_OFF1] = off1[i1]; rec[GI_L1] = P.in1[i1].l; rec[GI_MUL1] = P.in1[i1].mul;
        rec[GI_R0] = (int32_t)(re_w.size() / RE_W);
        for (const GrpH& gh : grps) {
            if (gh.i1 != (int)i1) continue;
            int32_t re[RE_W] = {0};
            re[RE_IO] = gh.io; re[RE_GRP] = gh.g; re[RE_WT0] = (int32_t)giwt_w.size();
            const int32_t* IOr = io_w.data() + gh.io * IO_W;
            for (int u0 = 0; u0 < P.in1[i1].mul; u0 += 4) {
                int found = -1;
                for (int b = 0; b < IOr[IO_NBLK]; ++b)
                    if (blk_w[IOr[IO_BLK] + b * BLK_W + B_GRP] == (gh.g | (u0 << 16))) found = b;
                giwt_w.push_back(IOr[IO_WTOFF] + 4 * blk_w[IOr[IO_BLK] + found * BLK_W + B_SUB0]);
            }
            re_w.insert(re_w.end(), re, re + RE_W);
            ++rec[GI_NR];
        }
        if (rec[GI_NR] > 0) gi_w.insert(gi_w.end(), rec, rec + GI_W);
    }
    P.nW = woff;
    if (P.nW == 0) {
        P.err = "o3tp: no path connects in1 x in2 to out";
        return false;
    }
    if (P.paths.size() >= 65536 || P.in2.size() > MAXP) {
        P.err = "o3tp: too many paths";
        return false;
    }
    std::vector<int32_t>& B = P.blob;
    B.assign(HDR_W, 0);
    B[H_NIO] = (int32_t)P.out.size(); B[H_D1] = P.D1; B[H_D2] = P.D2; B[H_DOUT] = P.Dout;
    B[H_IO] = (int32_t)B.size(); B.insert(B.end(), io_w.begin(), io_w.end());
    B[H_PATH] = (int32_t)B.size(); B.insert(B.end(), path_w.begin(), path_w.end());
    B[H_BLK] = (int32_t)B.size(); B.insert(B.end(), blk_w.begin(), blk_w.end());
    B[H_GRP] = (int32_t)B.size(); B.insert(B.end(), grp_w.begin(), grp_w.end());
    B[H_SUB] = (int32_t)B.size(); B.insert(B.end(), sub_w.begin(), sub_w.end());
    B[H_GI] = (int32_t)B.size(); B.insert(B.end(), gi_w.begin(), gi_w.end());
    B[H_NGI] = (int32_t)(gi_w.size() / GI_W);
    B[H_RE] = (int32_t)B.size(); B.insert(B.end(), re_w.begin(), re_w.end());
    B[H_GIWT] = (int32_t)B.size(); B.insert(B.end(), giwt_w.begin(), giwt_w.end());
    B[H_SPLIT] = split_ok ? 1 : 0; B[H_FMAX] = fmax;
    while (B.size() % 4) B.push_back(0);
    B[H_WORDS] = B[H_BASE] = (int32_t)B.size(); B[H_NW] = P.nW; B[H_FROW] = frow; B[H_GTMAX] = gtmax; B[H_MAXNP] = maxnp;
    B[H_NPATH] = (int32_t)P.paths.size(); B[H_NWP] = wsoff; B[H_NWT] = wtoff;
    return true;
}

// Forward work units of one tile of TE rows (TE a multiple of 32): (output irrep, channel chunk, 32-row group), spread
// over the NWARP warps by longest-processing-time-first on an instruction-count estimate.  Appended to the blob:
// [H_UNIT + w] .. [H_UNIT + w + 1] = range of warp w in the packed list (io | chunk << 8 | group << 16).
inline void schedule_forward(Plan& P, int TE) {
    std::vector<int32_t>& B = P.blob;
    struct U { int packed; double cost; };
    std::vector<U> us;
    for (int io = 0; io < B[H_NIO]; ++io) {
        const int32_t* IO = B.data() + B[H_IO] + io * IO_W;
        const int d = IO[IO_D], cw = IO[IO_CW];
        double cost = 40;
        for (int p = IO[IO_PBEG]; p < IO[IO_PEND]; ++p) {
            const int32_t* R = B.data() + B[H_PATH] + p * PATH_W;
            cost += 30 + R[P_D1] * d + R[P_MUL1] * (R[P_D1] + R[P_D1] * d + cw / 4 + d * cw + 4);
        }
        for (int q = 0; q < IO[IO_NQ]; ++q)
            for (int g = 0; g < TE / 32; ++g) us.push_back({io | (q << 8) | (g << 16), cost});
    }
    std::stable_sort(us.begin(), us.end(), [](const U& a, const U& b) { return a.cost > b.cost; });
    std::vector<std::vector<int>> per(NWARP);
    double load[NWARP] = {0};
    for (const U& u : us) {
        int w = 0;
        for (int k = 1; k < NWARP; ++k)
            if (load[k] < load[w]) w = k;
        per[w].push_back(u.packed);
        load[w] += u.cost;
    }
    B.resize(B[H_BASE]);  // drop an earlier schedule
    auto emit = [&](const std::vector<std::vector<int>>& lists) {
        const int base = (int)B.size();
        B.resize(base + NWARP + 1);
        int acc = 0;
        for (int w = 0; w < NWARP; ++w) {
            B[base + w] = acc;
            acc += (int)lists[w].size();
        }
        B[base + NWARP] = acc;
        for (int w = 0; w < NWARP; ++w) B.insert(B.end(), lists[w].begin(), lists[w].end());
        return base;
    };
    B[H_UNIT] = emit(per); B[H_TEF] = TE;
    // input-gradient kernel: units (in1 irrep, block of 4 channels, 32-row group), packed gi | block << 8 | group << 24
    us.clear();
    for (int gi = 0; gi < B[H_NGI]; ++gi) {
        const int32_t* GI = B.data() + B[H_GI] + gi * GI_W;
        const int d1 = 2 * GI[GI_L1] + 1;
        double cost = 40;
        for (int r = 0; r < GI[GI_NR]; ++r) {
            const int32_t* RE = B.data() + B[H_RE] + (GI[GI_R0] + r) * RE_W;
            const int32_t* IO = B.data() + B[H_IO] + RE[RE_IO] * IO_W;
            const int32_t* G = B.data() + B[H_GRP] + RE[RE_GRP] * GRP_W;
            cost += G[G_NP] * (IO[IO_MUL] * (5.0 * IO[IO_D] + 2) + 4.0 * d1 * IO[IO_D] + 30);
        }
        for (int ub = 0; ub * 4 < GI[GI_MUL1]; ++ub)
            for (int g = 0; g < TE_GIN / 32; ++g) us.push_back({gi | (ub << 8) | (g << 24), cost});
    }
    std::stable_sort(us.begin(), us.end(), [](const U& a, const U& b) { return a.cost > b.cost; });
    std::vector<std::vector<int>> per2(NWARP);
    double load2[NWARP] = {0};
    for (const U& u : us) {
        int w = 0;
        for (int k = 1; k < NWARP; ++k)
            if (load2[k] < load2[w]) w = k;
        per2[w].push_back(u.packed);
        load2[w] += u.cost;
    }
    B[H_GUNIT] = emit(per2);
    while (B.size() % 4) B.push_back(0);
    B[H_WORDS] = (int32_t)B.size();
}

// shared-memory floats of the split backward kernels
inline size_t gin_floats(const std::vector<int32_t>& B) {
    return (size_t)B[H_NWT] + (size_t)TE_GIN * (2 * (B[H_D1] | 1) + 2 * (B[H_D2] | 1) + (B[H_DOUT] | 1)) + 8;
}
inline size_t gw_floats(const std::vector<int32_t>& B) {
    return (size_t)std::max(B[H_FMAX], 16 * (32 * NWARP + 4)) + (size_t)B[H_GTMAX] +
           (size_t)TE_BWD * ((B[H_D1] | 1) + (B[H_D2] | 1) + (B[H_DOUT] | 1)) + 8;
}

// shared-memory floats of the tile programs (excluding the table), for a tile of TE rows
inline size_t fwd_floats(const std::vector<int32_t>& B, int TE) {
    return (size_t)B[H_NWP] + (size_t)TE * (2 * (B[H_D1] | 1) + 2 * (B[H_D2] | 1) + (B[H_DOUT] | 1)) + 8;
}
// backward (TE_BWD rows): resident transposed weights + gradient accumulators, x / gx / y / gy / g tiles, one round of
// features (4 rows per warp and path slot), the scaled transposed cotangent of one output irrep, and the scratch of
// the sliced weight-gradient partial sums (16 per thread)
inline size_t bwd_floats(const std::vector<int32_t>& B, bool resident_gw = true, bool dbuf = true) {
    const int nb = dbuf ? 2 : 1;
    return (size_t)B[H_NWT] + (resident_gw ? (size_t)B[H_NW] : 0) + (size_t)TE_BWD * ((nb + 1) * (B[H_D1] | 1) + (nb + 1) * (B[H_D2] | 1) + nb * (B[H_DOUT] | 1)) +
           (size_t)4 * B[H_MAXNP] * NWARP * B[H_FROW] + (size_t)B[H_GTMAX] + 16 * (32 * NWARP + 4) + 8;
}

}  // namespace o3
