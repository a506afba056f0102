// featurize_tiled.cu — fit path, register-tiled leg-grouped kernel for small 3-body grids.
//
// Same mathematics as the leg-grouped path of featurize.cu (the 3-body rows of
// angles.featurize_force_3b / featurize_energy_3b, representation/angles.py:17-286, factored
// per LEG GROUP instead of per triangle; unary trio of symmetry 2):
//   centre role     x_a += sum_j  u_aj dB_l(r_aj) (x) P_j,      P_j[m,n] = sum_{k != j} B_m(r_ak) B_n(r_jk)
//   energy          e   += sum_j  1/2 B_l(r_aj)  (x) P_j        (both orders of a pair fold onto one column)
//   neighbour role  x_a += -u_ia dB_l(r_ia) (x) P + B_l(r_ia) (x) Q,
//                   P[m,n] = sum_k B_m(r_ik) B_n(r_ak),   Q_c[m,n] = sum_k w_ak,c B_m(r_ik) dB_n(r_ak)
// but organised as small dense contractions with REGISTER tiles.  The previous kernel gave
// every lane one (m, n) cell and read both factors of every product from shared memory (one
// shared load per FMA: the L1/LSU pipe was 88 % busy, the FP64 pipe 10 %).  Here
//   * lane = (group g, n): a contraction round works on G = 32 / NA leg groups at once and a lane
//     keeps the [m] x {P, Qx, Qy, Qz} tile of its (group, n) in registers: one partner step is
//     NQ + 1 shared loads (the partner's A[m] and w, broadcast within the group; the lane's
//     (B_n, dB_n)) for 5 LM floating-point instructions;
//   * the [c][l][m] force accumulators of the lane's n stay in registers for the whole atom and
//     are folded into the compressed columns once per atom (lanes of different g are summed with
//     shuffles first);
//   * the legs (centre, neighbour) — needed 14 times each, by the centre and by every neighbour —
//     are evaluated once per frame by k_centre_legs into two small dense tables (96 B per list
//     entry, L2 resident), together with the position of every atom in its neighbour's row;
//   * the legs between two neighbours are evaluated in place by the lanes (two groups per pass)
//     from positions, dense by (basis index - first untrimmed index): no leg cache, no global
//     traffic beyond positions and list entries.
#include <algorithm>
#include <cstdlib>

#include "featurize_common.cuh"

namespace uf3b {

struct TiledGeom {
    int l0, n0, la, na;             // first untrimmed basis index and untrimmed extent: l/m legs, n leg
    int dim_m, dim_n, goff, col0;   // full grid extents of legs m, n; first bin; first feature column
    int nk_l, koff_l, poff_l;       // spline table of the l/m legs (doubles into knots3 / poly3)
    int nk_n, koff_n, poff_n;       // ... of the n leg
    double scale_l, scale_n;
    int n_knots3, n_poly3;          // table sizes (doubles), staged in shared memory per block
    int ps, cg, sl_shift;           // partner slots per group; groups per chunk; log2(lanes per group
                                    // in an evaluation pass)
    int off_warps, warp_bytes;      // per-warp regions behind the block's tables
    int off_own, off_grp, off_pos, off_zero, off_aw, off_vn;   // inside a warp's region: legs a -> e, legs centre(e) -> a',
                                            // ghost positions of the row, (A, w) records, dense n-leg records
    const double *legv, *legd, *epos;   // k_centre_legs tables
};

template <int LM, int NA>
struct TiledShape {
    static constexpr int G = 32 / NA;                 // leg groups per contraction round
    static constexpr int NQA = (LM + 1) / 2;          // 16-byte words holding A[LM]
    static constexpr int NQ = (LM + 3 + 1) / 2;       // ... holding A[LM], w[3]
    static constexpr unsigned AWB = 16u * NQ;         // bytes of an (A, w) record
    static constexpr unsigned VNB = 80;               // bytes of an n-leg record: the four non-zero (B_n, dB_n),
                                                      // then {first basis index - n0, pad}
};
constexpr unsigned OWN_REC = 96, POS_REC = 32;        // v[4] dv[4] u[3] pad; x y z pad
constexpr int TL_MAX_ROW = 32;                        // longest 3-body row the kernel takes
constexpr int TL_DEAD = -(1 << 20);                   // first basis index of a leg that contributes nothing

template <int I>
__device__ __forceinline__ double q_elem(const double2 *q) {
    return (I & 1) ? q[I >> 1].y : q[I >> 1].x;
}

// ---------------------------------------------------------------- k_centre_legs
// One thread per entry e of the 3-body row of centre a: the leg (a, entry) dense by
// (basis index - l0) — legv[p] = B[4], legd[p] = dB[4], unit vector a -> entry, and
// {parent atom of the entry, position of a's image in that atom's row (or -1)} packed in the
// last word; epos[p] = ghost position of the entry and {start, length} of its parent atom's row;
// p = off3[a] + e.  Same arithmetic as eval_dense_leg of featurize.cu.  The consumers then reach
// everything about a neighbour's row with ONE level of (coalesced) global loads instead of the
// chain list entry -> image -> position + offset.
__global__ void __launch_bounds__(256)
k_centre_legs(const BasisTab B, const FrameView f, const TiledGeom g, int sub_shift,
              double *__restrict__ legv, double *__restrict__ legd, double *__restrict__ epos) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int a = (int)(t >> sub_shift), e = (int)(t & ((1 << sub_shift) - 1));
    if (a >= f.n) return;
    const int row0 = __ldg(f.off3 + a), n3a = __ldg(f.cnt3 + a);
    if (e >= n3a) return;
    const int m = __ldg(f.idx3 + row0 + e);
    const int gimg = image_of(f, m);
    const int ci = m - gimg * f.n;
    const int apr = __ldg(f.img_inv + gimg) * f.n + a;
    const int rowi = __ldg(f.off3 + ci), ni = __ldg(f.cnt3 + ci);
    int lo = 0, hi = ni;                        // rows are sorted by supercell index
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(f.idx3 + rowi + mid) < apr) lo = mid + 1; else hi = mid;
    }
    const int qa = (lo < ni && __ldg(f.idx3 + rowi + lo) == apr) ? lo : -1;   // -1: one-ulp asymmetry
    int dummy;
    const Vec3 pa = real_position(f, a), pm = super_position(f, m, dummy);
    const double d = dist_rn(pa, pm);
    double2 *ov = reinterpret_cast<double2 *>(legv + 4 * (size_t)(row0 + e));
    double2 *od = reinterpret_cast<double2 *>(legd + 8 * (size_t)(row0 + e));
    ov[0] = ov[1] = od[0] = od[1] = make_double2(0.0, 0.0);
    double inv = 0.0;
    const double *kn = B.knots3 + g.koff_l;
    if (d >= __ldg(kn) && d <= __ldg(kn + g.nk_l - 1)) {        // angles.py:502-508
        double v[4], dv[4];
        const int idx = eval_leg(kn, g.nk_l, g.scale_l, B.poly3 + g.poff_l, d, B.lead3, B.trail3, v, dv);
        if (idx >= 0) {
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int x = idx + p - g.l0;
                if (x >= 0 && x < g.la) {
                    legv[4 * (size_t)(row0 + e) + x] = v[p];
                    legd[8 * (size_t)(row0 + e) + x] = dv[p];
                }
            }
        }
        inv = fast_rcp(d);
    }
    const long long tag = ((long long)qa << 32) | (unsigned)ci;
    od[2] = make_double2((pm.x - pa.x) * inv, (pm.y - pa.y) * inv);
    od[3] = make_double2((pm.z - pa.z) * inv, __longlong_as_double(tag));
    // the entry's ghost position (data/geometry.py:146-147) and the row of its parent atom
    double2 *op = reinterpret_cast<double2 *>(epos + 4 * (size_t)(row0 + e));
    op[0] = make_double2(pm.x, pm.y);
    op[1] = make_double2(pm.z, __longlong_as_double(((long long)ni << 32) | (unsigned)rowi));
}

// ---------------------------------------------------------------- leg between two neighbours
// Polynomial pieces (16 doubles) are staged in shared memory 144 bytes apart: with the natural
// 128-byte stride every lane whose distance falls into a different knot interval hit the same
// banks (72 % excess wavefronts on these loads, the largest shared-memory item of the kernel).
constexpr unsigned PIECE_B = 144;

struct NLegTab {
    unsigned knots_s, poly_s;       // shared-window addresses of the n-leg knots / pieces
    int nk;
    double scale;
};

// Values (and derivatives) of the four basis functions that are non-zero at d, and the first
// basis index, or TL_DEAD when the leg is outside its knot range (the reference drops the whole
// triangle, angles.py:502-508) or exactly on the first knot.  Same expressions as
// find_interval / eval_piece (spline.cuh) on tables staged in shared memory.
template <bool DERIV>
__device__ __forceinline__ int eval_n_leg(const NLegTab &T, double d, double (&v)[4], double (&dv)[4]) {
    const double t_lo = lds64(T.knots_s), t_hi = lds64(T.knots_s + 8u * (unsigned)(T.nk - 1));
    const double t_first = lds64(T.knots_s + 24u), t_last = lds64(T.knots_s + 8u * (unsigned)(T.nk - 4));
    if (!(d >= t_lo && d <= t_hi)) return TL_DEAD;
    if (!(d > t_first) || !(d <= t_last)) return TL_DEAD;
    int i = 3 + (int)((d - t_first) * T.scale);
    if (i > T.nk - 5) i = T.nk - 5;
    double ti = lds64(T.knots_s + 8u * (unsigned)i);
    const double ti1 = lds64(T.knots_s + 8u * (unsigned)i + 8u);
    if (!(ti < d && d <= ti1)) {
        int lo = 3, hi = T.nk - 4;          // invariant: t[lo] < d <= t[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (lds64(T.knots_s + 8u * (unsigned)mid) < d) lo = mid; else hi = mid;
        }
        i = lo;
        ti = lds64(T.knots_s + 8u * (unsigned)i);
    }
    const double u = d - ti;
    const unsigned piece = T.poly_s + PIECE_B * (unsigned)(i - 3);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double2 c01 = lds128(piece + 32u * q), c23 = lds128(piece + 32u * q + 16u);
        v[q] = ((c23.y * u + c23.x) * u + c01.y) * u + c01.x;
        if (DERIV) dv[q] = (3.0 * c23.y * u + 2.0 * c23.x) * u + c01.y;
    }
    return i - 3;
}

// n-leg record: the four non-zero (B_n, dB_n) and the window position rel = first basis index - n0
// (TL_DEAD for a leg that contributes nothing).  A contraction lane with untrimmed index n reads
// entry n - rel when that lies in [0, 4) and a zero word otherwise: trimmed basis functions have
// no lane, so no zero-filled dense array is ever written.
__device__ __forceinline__ void store_sparse_n(unsigned rec, int rel, const double (&v)[4], const double (&dv)[4]) {
#pragma unroll
    for (int p = 0; p < 4; ++p) sts128(rec + 16u * p, make_double2(v[p], dv[p]));
    asm volatile("st.shared.s32 [%0], %1;" :: "r"(rec + 64u), "r"(rel) : "memory");
}

// ---------------------------------------------------------------- contraction of one group
// P[m] = sum_k A_k[m] B_k[n];  FULL: also Q[c][m] = sum_k w_k[c] A_k[m] dB_k[n]  (lane's n), over the
// ns records of the group.  Software pipeline: the (A, w) record of step k + 1 and the window
// position of step k + 2 are loaded before the arithmetic of step k, the lane's (B_n, dB_n) of
// step k + 1 right behind it (two records of readable padding follow the last one).
template <int LM, int NA, bool FULL>
__device__ __forceinline__ void contract_group(unsigned aw, unsigned vn, int ns, int c_n, unsigned zero_s,
                                               double (&P)[LM], double (&Q)[3][LM]) {
    using S = TiledShape<LM, NA>;
    constexpr int NL = FULL ? S::NQ : S::NQA;
#pragma unroll
    for (int m = 0; m < LM; ++m) { P[m] = 0.0; Q[0][m] = Q[1][m] = Q[2][m] = 0.0; }
    if (ns <= 0) return;
    auto pick = [&](unsigned rec, int rel) {        // address of the lane's entry of a record
        const int qq = c_n - rel;
        return (unsigned)qq < 4u ? rec + 16u * (unsigned)qq : zero_s;
    };
    double2 q[S::NQ], b;
#pragma unroll
    for (int i = 0; i < NL; ++i) q[i] = lds128(aw + 16u * i);
    {
        const unsigned ad = pick(vn, lds32(vn + 64u));
        if (FULL) b = lds128(ad); else b = make_double2(lds64(ad), 0.0);
    }
    int rel1 = lds32(vn + S::VNB + 64u);            // window position of the next record
#pragma unroll 2
    for (int s = 0; s < ns; ++s) {
        aw += S::AWB;
        vn += S::VNB;
        double2 qn[S::NQ], bn;
#pragma unroll
        for (int i = 0; i < NL; ++i) qn[i] = lds128(aw + 16u * i);
        const int rel2 = lds32(vn + S::VNB + 64u);
        {
            const unsigned ad = pick(vn, rel1);
            if (FULL) bn = lds128(ad); else bn = make_double2(lds64(ad), 0.0);
        }
        double A[LM];
        A[0] = q_elem<0>(q);
        if constexpr (LM > 1) A[1] = q_elem<1>(q);
        if constexpr (LM > 2) A[2] = q_elem<2>(q);
        if constexpr (LM > 3) A[3] = q_elem<3>(q);
        if (FULL) {
            const double w0 = q_elem<LM>(q), w1 = q_elem<LM + 1>(q), w2 = q_elem<LM + 2>(q);
#pragma unroll
            for (int m = 0; m < LM; ++m) {
                const double t = A[m] * b.y;
                P[m] = fma(A[m], b.x, P[m]);
                Q[0][m] = fma(w0, t, Q[0][m]);
                Q[1][m] = fma(w1, t, Q[1][m]);
                Q[2][m] = fma(w2, t, Q[2][m]);
            }
        } else {
#pragma unroll
            for (int m = 0; m < LM; ++m) P[m] = fma(A[m], b.x, P[m]);
        }
#pragma unroll
        for (int i = 0; i < NL; ++i) q[i] = qn[i];
        b = bn;
        rel1 = rel2;
    }
}

// index of the unordered pair {l, m} in the folded tile: (l, m, n) and (m, l, n) share a column
// (symmetry 2), so the accumulators are kept for l <= m only
__host__ __device__ constexpr int sym_idx(int l, int m, int LM) {
    return l <= m ? l * LM - l * (l - 1) / 2 + (m - l) : m * LM - m * (m - 1) / 2 + (l - m);
}

// ---------------------------------------------------------------- the kernel
// Block = W independent warps (W chosen by the host from the shared-memory budget), warp = one
// atom at a time.  Per-warp shared memory: accumulators [col0][e, fx, fy, fz] of the composition
// and pair columns (layout of featurize.cu, shared with two_body_rows), three small tables of the
// atom's own row (legs a -> e, legs centre(e) -> a', ghost positions), and the chunk buffers of
// cg groups x ps partner slots (slot = position in the centre's row; the slot of the group's own
// atom stays zero).  The 3-body columns never pass through shared memory: the folded
// [c][{l, m}] tile of the lane's n lives in registers and is stored straight into the rows.
template <int LM, int NA>
__global__ void __launch_bounds__(128, 4)
k_featurize_tiled(const BasisTab B, const FrameView f, const TiledGeom tg, double *__restrict__ xf, long long ld,
                  double *__restrict__ partials, int want_e_, int want_f_) {
    using S = TiledShape<LM, NA>;
    constexpr int NS = LM * (LM + 1) / 2;
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int gw = blockIdx.x * nw + warp, n_gw = gridDim.x * nw;
    const int F = B.n_feats;
    const bool want_e = want_e_ != 0, want_f = want_f_ != 0;

    // ---- block: n-leg spline table (knots, then pieces PIECE_B apart)
    const int n_kn = (tg.nk_n + 1) & ~1;
    {
        double *tab = reinterpret_cast<double *>(smem);
        const double *kn = B.knots3 + tg.koff_n, *po = B.poly3 + tg.poff_n;
        const int n_po = 16 * (tg.nk_n - 7);
        for (int k = threadIdx.x; k < tg.nk_n; k += blockDim.x) tab[k] = __ldg(kn + k);
        for (int k = threadIdx.x; k < n_po; k += blockDim.x)
            tab[n_kn + (k >> 4) * (int)(PIECE_B / 8) + (k & 15)] = __ldg(po + k);
    }
    __syncthreads();
    const unsigned smem_s = pin(smem_addr(smem));
    NLegTab nt;
    nt.knots_s = smem_s;
    nt.poly_s = smem_s + 8u * (unsigned)n_kn;
    nt.nk = tg.nk_n;
    nt.scale = tg.scale_n;

    // ---- warp regions
    unsigned char *mine = smem + tg.off_warps + (size_t)warp * (size_t)tg.warp_bytes;
    const unsigned mine_s = pin(smem_s + (unsigned)tg.off_warps + (unsigned)warp * (unsigned)tg.warp_bytes);
    double *acc = reinterpret_cast<double *>(mine);              // [col0][e, fx, fy, fz]
    PairRec *prec = reinterpret_cast<PairRec *>(mine + tg.off_aw);      // pair pass scratch (aliases the chunk buffers)
    const int c_g = lane / NA;                                           // contraction role of the lane: group,
    const int c_n = (lane - c_g * NA < tg.na) ? lane - c_g * NA : -(1 << 24);   // n (lanes outside the window pick nothing)
    const bool c_on = c_g < S::G;
    const double half_e = want_e ? 0.5 : 0.0;
    const int n_acc = 4 * tg.col0;

    for (int k = lane; k < n_acc; k += 32) acc[k] = 0.0;
    if (lane == 0) *reinterpret_cast<double2 *>(mine + tg.off_zero) = make_double2(0.0, 0.0);
    double er[NS];              // energy tile {l, m} of the lane's (g, n), kept over all atoms of the warp
#pragma unroll
    for (int s = 0; s < NS; ++s) er[s] = 0.0;
    __syncwarp();

    for (int a = gw; a < f.n; a += n_gw) {
        {
            const int sa = __ldg(f.spec + a);
            const Vec3 pa = real_position(f, a);
            if (lane == 0) acc[4 * sa] += 1.0;      // composition column n_el (composition.py:96-111)
            __syncwarp();
            // -------------------------------------------- 2-body (bspline.py:810-895)
            two_body_rows(B, f, a, sa, pa, acc, prec, lane);
        }

        // ------------------------------------------------ 3-body
        const unsigned own_s = mine_s + (unsigned)tg.off_own, grp_s = mine_s + (unsigned)tg.off_grp;
        const unsigned pos_s = mine_s + (unsigned)tg.off_pos;
        const unsigned aw_s = mine_s + (unsigned)tg.off_aw, vn_s = mine_s + (unsigned)tg.off_vn;
        const unsigned zero_s = mine_s + (unsigned)tg.off_zero;
        const int ps = tg.ps, cg = tg.cg;
        const int gpp = 32 >> tg.sl_shift;                               // groups per evaluation pass
        const int e_gg = lane >> tg.sl_shift, e_k = lane & ((1 << tg.sl_shift) - 1);
        const int row0 = __ldg(f.off3 + a), n3a = __ldg(f.cnt3 + a);
        double fr[3][NS];           // force tile [c][{l, m}] of the lane's (g, n)
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int s = 0; s < NS; ++s) fr[c][s] = 0.0;

        // the atom's own row from the k_centre_legs tables: legs (a, e), ghost positions, and for
        // the neighbour role the row of every neighbour's parent atom and that centre's leg to a'
        int my_rowi = 0, my_ni = 0, my_qa = -1;
        if (lane < n3a) {
            const size_t p = (size_t)(row0 + lane);
            const double2 *gv = reinterpret_cast<const double2 *>(tg.legv + 4 * p);
            const double2 *gd = reinterpret_cast<const double2 *>(tg.legd + 8 * p);
            const double2 *gp = reinterpret_cast<const double2 *>(tg.epos + 4 * p);
            const double2 v01 = __ldg(gv), v23 = __ldg(gv + 1), d01 = __ldg(gd), d23 = __ldg(gd + 1);
            const double2 u01 = __ldg(gd + 2), u2t = __ldg(gd + 3), pxy = __ldg(gp), pzt = __ldg(gp + 1);
            my_qa = (int)(__double_as_longlong(u2t.y) >> 32);
            const long long rt = __double_as_longlong(pzt.y);
            my_rowi = (int)(rt & 0xffffffffll);
            my_ni = (int)(rt >> 32);
            const unsigned o = own_s + OWN_REC * (unsigned)lane;
            sts128(o, v01); sts128(o + 16, v23); sts128(o + 32, d01); sts128(o + 48, d23);
            sts128(o + 64, u01); sts64(o + 80, u2t.x);
            const unsigned po = pos_s + POS_REC * (unsigned)lane;
            sts128(po, pxy);
            sts64(po + 16, pzt.x);
            if (want_f && my_qa >= 0) {
                const size_t p2 = (size_t)(my_rowi + my_qa);
                const double2 *hv = reinterpret_cast<const double2 *>(tg.legv + 4 * p2);
                const double2 *hd = reinterpret_cast<const double2 *>(tg.legd + 8 * p2);
                const double2 a0 = __ldg(hv), a1 = __ldg(hv + 1), b0 = __ldg(hd), b1 = __ldg(hd + 1);
                const double2 c0_ = __ldg(hd + 2), c1_ = __ldg(hd + 3);
                const unsigned g_ = grp_s + OWN_REC * (unsigned)lane;
                sts128(g_, a0); sts128(g_ + 16, a1); sts128(g_ + 32, b0); sts128(g_ + 48, b1);
                sts128(g_ + 64, c0_); sts64(g_ + 80, c1_.x);
            }
        }
        __syncwarp();

        // ---- (ii) `a` as a neighbour of the centre i named by entry e of its row
        if (want_f) {
            for (int c0 = 0; c0 < n3a; c0 += cg) {
                const int ng = min(cg, n3a - c0);
                for (int gb = 0; gb < ng; gb += gpp) {          // evaluation passes: legs (a', k) and A = B_m(r_ik)
                    const int gi = gb + e_gg;
                    const int e = min(c0 + gi, TL_MAX_ROW - 1);
                    const int rowi = __shfl_sync(FULL, my_rowi, e), ni = __shfl_sync(FULL, my_ni, e);
                    const int qa = __shfl_sync(FULL, my_qa, e);
                    if (gi < ng && qa >= 0 && e_k < ni) {
                        const double2 *ga = reinterpret_cast<const double2 *>(tg.epos + 4 * (size_t)(rowi + qa));
                        const double2 *gk = reinterpret_cast<const double2 *>(tg.epos + 4 * (size_t)(rowi + e_k));
                        const double2 *gv = reinterpret_cast<const double2 *>(tg.legv + 4 * (size_t)(rowi + e_k));
                        const double2 axy = __ldg(ga), azt = __ldg(ga + 1), kxy = __ldg(gk), kzt = __ldg(gk + 1);
                        double2 q[S::NQ];
#pragma unroll
                        for (int i = 0; i < S::NQA; ++i) q[i] = __ldg(gv + i);
                        const Vec3 pap = {axy.x, axy.y, azt.x}, pk = {kxy.x, kxy.y, kzt.x};
                        const double d = dist_rn(pap, pk);
                        double v[4], dv[4];
                        const int idx = eval_n_leg<true>(nt, d, v, dv);
                        const unsigned slot = (unsigned)(gi * ps + e_k);
                        store_sparse_n(vn_s + S::VNB * slot, idx < 0 ? TL_DEAD : idx - tg.n0, v, dv);
                        const double inv = idx < 0 ? 0.0 : fast_rcp(d);
                        const double w[3] = {(pk.x - pap.x) * inv, (pk.y - pap.y) * inv, (pk.z - pap.z) * inv};
                        // (A[LM], w[3]) packed behind each other
                        double rec[2 * S::NQ];
#pragma unroll
                        for (int i = 0; i < S::NQA; ++i) { rec[2 * i] = q[i].x; rec[2 * i + 1] = q[i].y; }
                        rec[LM] = w[0]; rec[LM + 1] = w[1]; rec[LM + 2] = w[2];
                        if (LM + 3 < 2 * S::NQ) rec[2 * S::NQ - 1] = 0.0;
                        const unsigned ar = aw_s + S::AWB * slot;
#pragma unroll
                        for (int i = 0; i < S::NQ; ++i) sts128(ar + 16u * i, make_double2(rec[2 * i], rec[2 * i + 1]));
                    }
                }
                __syncwarp();
                for (int r0 = 0; r0 < ng; r0 += S::G) {         // contraction rounds
                    const int gi = r0 + c_g;
                    const int e = min(c0 + gi, TL_MAX_ROW - 1);
                    const int ni = __shfl_sync(FULL, my_ni, e), qa = __shfl_sync(FULL, my_qa, e);
                    const bool on = c_on && gi < ng && qa >= 0;
                    double P[LM], Q[3][LM];
                    contract_group<LM, NA, true>(aw_s + S::AWB * (unsigned)(gi * ps), vn_s + S::VNB * (unsigned)(gi * ps),
                                                 on ? ni : 0, c_n, zero_s, P, Q);
                    if (on) {       // the centre's leg to a': x_a += -u dB_l (x) P + B_l (x) Q
                        const unsigned ge = grp_s + OWN_REC * (unsigned)e;
                        double2 q[6];
#pragma unroll
                        for (int i = 0; i < 5; ++i) q[i] = lds128(ge + 16u * i);
                        q[5].x = lds64(ge + 80);
                        const double vl[4] = {q[0].x, q[0].y, q[1].x, q[1].y}, dl[4] = {q[2].x, q[2].y, q[3].x, q[3].y};
                        const double u[3] = {q[4].x, q[4].y, q[5].x};
#pragma unroll
                        for (int l = 0; l < LM; ++l) {
                            const double n0_ = -u[0] * dl[l], n1_ = -u[1] * dl[l], n2_ = -u[2] * dl[l];
#pragma unroll
                            for (int m = 0; m < LM; ++m) {
                                constexpr int dummy_ = 0; (void)dummy_;
                                const int s = sym_idx(l, m, LM);
                                fr[0][s] = fma(n0_, P[m], fma(vl[l], Q[0][m], fr[0][s]));
                                fr[1][s] = fma(n1_, P[m], fma(vl[l], Q[1][m], fr[1][s]));
                                fr[2][s] = fma(n2_, P[m], fma(vl[l], Q[2][m], fr[2][s]));
                            }
                        }
                    }
                }
                __syncwarp();
            }
        }

        // ---- (i) `a` as the centre: group j = its leg to neighbour j, partners k != j
        if (n3a > 1) {
            for (int c0 = 0; c0 < n3a; c0 += cg) {
                const int ng = min(cg, n3a - c0);
                for (int gb = 0; gb < ng; gb += gpp) {          // evaluation passes: legs (j, k), values only
                    const int gi = gb + e_gg, j = c0 + gi;
                    if (gi < ng && e_k < n3a) {
                        const unsigned pj_ = pos_s + POS_REC * (unsigned)j, pk_ = pos_s + POS_REC * (unsigned)e_k;
                        const double2 jxy = lds128(pj_), kxy = lds128(pk_);
                        const Vec3 pj = {jxy.x, jxy.y, lds64(pj_ + 16)}, pk = {kxy.x, kxy.y, lds64(pk_ + 16)};
                        double v[4], dv[4] = {0.0, 0.0, 0.0, 0.0};
                        const int idx = eval_n_leg<false>(nt, dist_rn(pj, pk), v, dv);
                        const unsigned slot = (unsigned)(gi * ps + e_k);
                        store_sparse_n(vn_s + S::VNB * slot, idx < 0 ? TL_DEAD : idx - tg.n0, v, dv);
                        const unsigned ok = own_s + OWN_REC * (unsigned)e_k, ar = aw_s + S::AWB * slot;
#pragma unroll
                        for (int i = 0; i < S::NQA; ++i) sts128(ar + 16u * i, lds128(ok + 16u * i));
                    }
                }
                __syncwarp();
                for (int r0 = 0; r0 < ng; r0 += S::G) {         // contraction rounds
                    const int gi = r0 + c_g, j = c0 + gi;
                    const bool on = c_on && gi < ng;
                    double P[LM], Q[3][LM];
                    contract_group<LM, NA, false>(aw_s + S::AWB * (unsigned)(gi * ps), vn_s + S::VNB * (unsigned)(gi * ps),
                                                  on ? n3a : 0, c_n, zero_s, P, Q);
                    if (on) {       // x_a += u_aj dB_l (x) P_j,  e += 1/2 B_l (x) P_j
                        const unsigned oj = own_s + OWN_REC * (unsigned)j;
                        double2 q[6];
#pragma unroll
                        for (int i = 0; i < 5; ++i) q[i] = lds128(oj + 16u * i);
                        q[5].x = lds64(oj + 80);
                        const double vl[4] = {q[0].x, q[0].y, q[1].x, q[1].y}, dl[4] = {q[2].x, q[2].y, q[3].x, q[3].y};
                        const double u[3] = {q[4].x, q[4].y, q[5].x};
#pragma unroll
                        for (int l = 0; l < LM; ++l) {
                            const double hv = half_e * vl[l];
#pragma unroll
                            for (int m = 0; m < LM; ++m) {
                                const int s = sym_idx(l, m, LM);
                                er[s] = fma(hv, P[m], er[s]);
                                const double dP = dl[l] * P[m];
                                fr[0][s] = fma(u[0], dP, fr[0][s]);
                                fr[1][s] = fma(u[1], dP, fr[1][s]);
                                fr[2][s] = fma(u[2], dP, fr[2][s]);
                            }
                        }
                    }
                }
                __syncwarp();
            }
        }

        // ------------------------------------------------ rows fx_a, fy_a, fz_a
        if (want_f) {
            // 3-body columns: lanes of different g are summed, lane n < na then owns every bin of
            // its n — (l, m, n) and (m, l, n) share a column — and stores it straight into the rows
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    double t = fr[c][s];
#pragma unroll
                    for (int g = 1; g < S::G; ++g) t += __shfl_down_sync(FULL, fr[c][s], g * NA);
                    fr[c][s] = t;
                }
            if (lane < tg.na) {
#pragma unroll
                for (int l = 0; l < LM; ++l)
#pragma unroll
                    for (int m = l; m < LM; ++m) {
                        if (m < tg.la) {
                            const int col = __ldg(B.bin_col + tg.goff + ((tg.l0 + l) * tg.dim_m + tg.l0 + m) * tg.dim_n + tg.n0 + lane);
                            if (col >= 0) {
                                double *dst = xf + (long long)a * ld + tg.col0 + col;
                                const int s = sym_idx(l, m, LM);
                                __stcs(dst, fr[0][s]);
                                __stcs(dst + (long long)f.n * ld, fr[1][s]);
                                __stcs(dst + 2 * (long long)f.n * ld, fr[2][s]);
                            }
                        }
                    }
            }
            // composition and pair columns
            for (int col = lane; col < tg.col0; col += 32) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    __stcs(xf + ((long long)c * f.n + a) * ld + col, acc[4 * col + 1 + c]);   // written once: streaming
                    acc[4 * col + 1 + c] = 0.0;
                }
            }
        }
        __syncwarp();
    }
    if (want_e) {
        double *mine_p = partials + (size_t)gw * F;
#pragma unroll
        for (int l = 0; l < LM; ++l)
#pragma unroll
            for (int m = l; m < LM; ++m) {
                const int s = sym_idx(l, m, LM);
                double t = er[s];
#pragma unroll
                for (int g = 1; g < S::G; ++g) t += __shfl_down_sync(FULL, er[s], g * NA);
                if (lane < tg.na && m < tg.la) {
                    const int col = __ldg(B.bin_col + tg.goff + ((tg.l0 + l) * tg.dim_m + tg.l0 + m) * tg.dim_n + tg.n0 + lane);
                    if (col >= 0) mine_p[tg.col0 + col] = t;
                }
            }
        for (int col = lane; col < tg.col0; col += 32) mine_p[col] = acc[4 * col];
    }
}

// ---------------------------------------------------------------- host side
template <int LM, int NA>
static int launch_tiled(uf3b_basis *basis, const uf3b_nlist *nl, TiledGeom tg, double *x_energy, double *x_forces,
                        int64_t ld, cudaStream_t stream) {
    using S = TiledShape<LM, NA>;
    const int F = basis->n_feats, n = (int)nl->n;
    const bool e_dev = x_energy && is_device_pointer(x_energy);
    const bool f_dev = x_forces && is_device_pointer(x_forces);
    int dev = 0, smem_max = 0, smem_sm = 0;
    UF3B_CUDA(cudaGetDevice(&dev));
    UF3B_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    UF3B_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));

    const int max3 = std::max(nl->max3, 2);
    tg.ps = max3;                   // one slot per position of the longest row
    tg.sl_shift = tg.ps <= 16 ? 4 : 5;
    // a chunk of cg groups holds whole contraction rounds (G groups); the larger candidate also holds
    // whole evaluation passes (32 >> sl_shift groups) and is taken when it does not cost resident warps
    const int cg_env = getenv("UF3B_TILED_CG") ? atoi(getenv("UF3B_TILED_CG")) : 0;
    const int w_env = getenv("UF3B_TILED_WARPS") ? atoi(getenv("UF3B_TILED_WARPS")) : 0;
    const int gpp = 32 >> tg.sl_shift;
    const int rows = (max3 + 1) & ~1;
    const size_t tab_bytes = 8 * (size_t)((tg.nk_n + 1) & ~1) + (size_t)PIECE_B * (tg.nk_n - 7);
    tg.off_warps = (int)((tab_bytes + 15) & ~size_t(15));
    tg.off_own = (int)((4 * (size_t)tg.col0 * sizeof(double) + 15) & ~size_t(15));
    tg.off_grp = tg.off_own + rows * (int)OWN_REC;
    tg.off_pos = tg.off_grp + rows * (int)OWN_REC;
    tg.off_zero = tg.off_pos + rows * (int)POS_REC;
    tg.off_aw = tg.off_zero + 16;
    auto layout = [&](int cg, int &warps, int &blocks) {
        tg.cg = cg;
        // two records of readable padding behind each array (the contraction loads ahead)
        const size_t aw_bytes = ((size_t)cg * tg.ps + 2) * S::AWB;
        tg.off_vn = tg.off_aw + (int)aw_bytes;
        size_t chunk_bytes = aw_bytes + ((size_t)cg * tg.ps + 2) * S::VNB;
        chunk_bytes = std::max(chunk_bytes, 32 * sizeof(PairRec));     // the pair pass borrows the chunk buffers
        tg.warp_bytes = (int)((tg.off_aw + chunk_bytes + 15) & ~size_t(15));
        // warps per block x blocks per SM: most resident warps within the register budget (16 warps
        // at 128 registers) and the shared memory of an SM (1 KB reserved per block)
        warps = blocks = 0;
        for (int b = 2; b <= 8; ++b)
            for (int w = 4; w >= 1; --w) {
                if (w * b > 16) continue;
                const size_t blk = (size_t)tg.off_warps + (size_t)w * tg.warp_bytes;
                if (blk > (size_t)smem_max || (size_t)b * (blk + 1024) > (size_t)smem_sm) continue;
                if (w * b > warps * blocks || (w * b == warps * blocks && w > warps)) { warps = w; blocks = b; }
            }
        if (warps == 0 && (size_t)tg.off_warps + tg.warp_bytes <= (size_t)smem_max) {     // one block per SM
            warps = (int)std::min<size_t>(4, ((size_t)smem_max - tg.off_warps) / tg.warp_bytes);
            blocks = 1;
        }
    };
    int warps = 0, blocks = 0;
    if (cg_env > 0) {
        layout(cg_env, warps, blocks);
    } else {
        const int big = S::G % gpp == 0 ? S::G : S::G * gpp;
        int w2 = 0, b2 = 0;
        layout(big, warps, blocks);
        if (big != S::G) {
            layout(S::G, w2, b2);
            if (w2 * b2 > warps * blocks) { warps = w2; blocks = b2; } else layout(big, warps, blocks);
        }
    }
    if (warps < 1) return 1;            // does not fit: the caller takes another path
    if (w_env > 0 && w_env <= 4 && (size_t)tg.off_warps + (size_t)w_env * tg.warp_bytes <= (size_t)smem_max) warps = w_env;
    auto kernel = k_featurize_tiled<LM, NA>;
    const size_t smem = (size_t)tg.off_warps + (size_t)warps * tg.warp_bytes;
    UF3B_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    UF3B_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, warps * 32, smem));
    if (per_sm < 1) per_sm = 1;
    int grid = std::max(1, sm_count() * per_sm / std::max(1, basis->frames_in_flight));
    grid = std::min(grid, (n + warps - 1) / warps);
    const int n_gw = grid * warps;

    // tables of k_centre_legs: one slot per list entry (rows live in claim regions of idx3)
    const size_t entries = nl->idx3.cap;
    UF3B_CUDA(basis->legv.reserve(4 * entries));
    UF3B_CUDA(basis->legd.reserve(8 * entries));
    UF3B_CUDA(basis->epos.reserve(4 * entries));
    tg.legv = basis->legv.p;
    tg.legd = basis->legd.p;
    tg.epos = basis->epos.p;

    double *d_xf = x_forces;
    long long d_ld = ld;
    if (x_forces && !f_dev) {
        UF3B_CUDA(basis->stage.reserve((size_t)3 * n * F));
        d_xf = basis->stage.p;
        d_ld = F;
    }
    double *d_xe = x_energy;
    if (x_energy) {
        UF3B_CUDA(basis->partials.reserve((size_t)(n_gw + ER_SPLIT) * F));
        if (!e_dev) {
            UF3B_CUDA(basis->stage_e.reserve(F));
            d_xe = basis->stage_e.p;
        }
    }
    const FrameView view = nl->view();
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (g_timing) {
        UF3B_CUDA(cudaEventCreate(&ev0));
        UF3B_CUDA(cudaEventCreate(&ev1));
        UF3B_CUDA(cudaEventRecord(ev0, stream));
    }
    const int sub_shift = max3 <= 16 ? 4 : 5;
    const long long threads = (long long)n << sub_shift;
    UF3B_LAUNCH(k_centre_legs, (unsigned)((threads + 255) / 256), 256, 0, stream, basis->tab, view, tg, sub_shift,
                basis->legv.p, basis->legd.p, basis->epos.p);
    UF3B_LAUNCH(kernel, grid, warps * 32, smem, stream, basis->tab, view, tg, d_xf, d_ld, basis->partials.p,
                x_energy ? 1 : 0, x_forces ? 1 : 0);
    if (g_timing) UF3B_CUDA(cudaEventRecord(ev1, stream));
    if (x_energy)
        if (int rc = launch_energy_row(basis->partials.p, n_gw, F, d_xe, stream)) return rc;
    return finish_featurize(basis, x_energy, x_forces, ld, d_xe, d_xf, F, n, e_dev, f_dev, stream, ev0, ev1);
}

// Takes the frame if the basis fits the tiled kernel: unary trio of symmetry 2 with unit folding
// weights, untrimmed grid la x la x na with la <= 4, na <= 10, rows of at most 32 entries.
// Returns 1 when it does not apply (the caller goes on to the other paths); error codes are <= 0.
int featurize_tiled(uf3b_basis *basis, const uf3b_nlist *nl, double *x_energy, double *x_forces, int64_t ld,
                    cudaStream_t stream) {
    const BasisTab &T = basis->tab;
    if (T.n_trios != 1 || !T.unit_weights || basis->no_tile || getenv("UF3B_NO_TILED") || getenv("UF3B_NO_LEGS")
        || getenv("UF3B_PLANES"))
        return 1;
    if (basis->h_trio_sym[0] != 2 || nl->max3 > TL_MAX_ROW) return 1;
    const int lead = T.lead3, trail = T.trail3;
    const int L = basis->h_trio_dims[0], M = basis->h_trio_dims[1], N = basis->h_trio_dims[2];
    TiledGeom tg = {};
    tg.l0 = tg.n0 = lead;
    tg.la = L - lead - trail;
    tg.na = N - lead - trail;
    if (L != M || tg.la < 1 || tg.na < 1 || tg.la > 4 || tg.na > 10) return 1;
    // the l and m legs must share their knots (symmetry 2 of a BSplineBasis guarantees it)
    {
        const double *k0 = basis->h_knots3.data() + basis->h_trio_koff[0], *k1 = basis->h_knots3.data() + basis->h_trio_koff[1];
        for (int k = 0; k < L + 4; ++k)
            if (k0[k] != k1[k]) return 1;
    }
    tg.dim_m = M;
    tg.dim_n = N;
    tg.goff = basis->h_trio_goff[0];
    tg.col0 = basis->h_trio_col[0];
    tg.nk_l = L + 4; tg.koff_l = basis->h_trio_koff[0]; tg.poff_l = basis->h_trio_poff[0];
    tg.nk_n = N + 4; tg.koff_n = basis->h_trio_koff[2]; tg.poff_n = basis->h_trio_poff[2];
    tg.scale_l = basis->h_trio_scale[0];
    tg.scale_n = basis->h_trio_scale[2];
    tg.n_knots3 = basis->n_knots3;
    tg.n_poly3 = basis->n_poly3;
    // the kernel stores the 3-body columns from the folded {l, m} tile: they must be the last
    // columns of the row, (l, m, n) and (m, l, n) must share a column, and the cells with l <= m
    // must reach every column exactly once
    {
        int n_cols3 = 0;
        for (int b = 0; b < L * M * N; ++b) n_cols3 = std::max(n_cols3, basis->h_bin_col[tg.goff + b] + 1);
        if (tg.col0 + n_cols3 != basis->n_feats) return 1;
        std::vector<int> seen(n_cols3, 0);
        for (int l = 0; l < L; ++l)
            for (int m = 0; m < M; ++m)
                for (int n = 0; n < N; ++n) {
                    const int c = basis->h_bin_col[tg.goff + (l * M + m) * N + n];
                    if (c != basis->h_bin_col[tg.goff + (m * M + l) * N + n]) return 1;
                    const bool inside = l >= lead && l < lead + tg.la && m >= lead && m < lead + tg.la
                                        && n >= lead && n < lead + tg.na;
                    if (c >= 0 && !inside) return 1;
                    if (c >= 0 && l <= m) seen[c]++;
                }
        for (int c = 0; c < n_cols3; ++c)
            if (seen[c] != 1) return 1;
    }
    if (tg.la <= 2 && tg.na <= 7) return launch_tiled<2, 7>(basis, nl, tg, x_energy, x_forces, ld, stream);
    if (tg.la <= 3 && tg.na <= 9) return launch_tiled<3, 9>(basis, nl, tg, x_energy, x_forces, ld, stream);
    return launch_tiled<4, 10>(basis, nl, tg, x_energy, x_forces, ld, stream);
}

}  // namespace uf3b
