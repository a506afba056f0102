// featurize_tiled.cu — fit path, register-tiled leg-grouped kernels for small 3-body grids.
//
// Same mathematics as the leg-grouped path of featurize.cu (the 3-body rows of
// angles.featurize_force_3b / featurize_energy_3b, representation/angles.py:17-286, factored
// per LEG GROUP instead of per triangle; unary trio of symmetry 2).  With the PLANE of a group
//   P_(i,j)[m,n] = sum_{k != j} B_m(r_ik) B_n(r_jk)            (centre i, its neighbour j, partners k)
// the rows of atom a are
//   neighbour role  x_a += sum_{i in row(a)}  -u_ia dB_l(r_ia) (x) P_(i,a') + B_l(r_ia) (x) Q,
//                   Q_c[m,n] = sum_k w_a'k,c B_m(r_ik) dB_n(r_a'k)
//   centre role     x_a += sum_{j in row(a)}   u_aj dB_l(r_aj) (x) P_(a,j)
//   energy          e   += sum_{j in row(a)}   1/2 B_l(r_aj)  (x) P_(a,j)   (both orders of a pair fold onto one column)
// The plane P_(i,j) is needed twice: by atom j in its neighbour role and by atom i in its centre
// role.  It is computed ONCE — by j, whose neighbour role needs the legs (j, k) anyway — and left in
// an L2-resident table (224 B per list entry); the centre role of every atom is then 14 plane reads
// and outer products instead of a second round of leg evaluations and contractions.
//
//   k_centre_legs   one thread per 3-body list entry: the leg (centre, entry) dense by basis index,
//                   the entry's ghost position, the row of its parent atom and the position of the
//                   centre's image in that row — so the consumers reach everything about a
//                   neighbour's row with ONE level of coalesced loads.
//   k_rows_nbr      warp = atom, neighbour role.  Lanes evaluate the legs (a', k) of two groups per
//                   pass into sparse records in shared memory; then lane = (group g, n) contracts
//                   G = 32 / NA groups per round with the [m] x {P, Qx, Qy, Qz} tile of its (g, n) in
//                   registers (one partner step = NQ + 2 shared loads for 5 LM FP64 instructions —
//                   the previous kernel read both factors of EVERY product from shared memory and
//                   sat at 88 % of the L1/LSU pipe with the FP64 pipe at 10 %), stores P to the
//                   plane table and adds the outer products to the folded [c][{l, m}] force tile of
//                   its n, which lives in registers for the whole atom and is left, 3 NS doubles per
//                   lane, in an L2-resident tile buffer.
//   k_rows_ctr      warp = atom.  Three TMA bulk copies (cp.async.bulk + per-warp mbarrier) stage the
//                   atom's row tables — leg values, derivatives, planes — while the lanes gather the
//                   pair partners; centre role from the planes, pair rows + composition columns; the
//                   force tile starts from the one k_rows_nbr left, the rows are written ONCE
//                   (streaming stores), the consumed tile and plane lines are dropped from the L2, and
//                   the energy-row partials are produced.
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
    int ps, cg, sl_shift;           // partner slots per group; groups per chunk; log2(lanes per group
                                    // in an evaluation pass)
    int off_warps, warp_bytes;      // k_rows_nbr: per-warp regions behind the block's spline table
    int off_zero, off_aw, off_vn;   // inside a warp's region: [legs centre(e) -> a'] zero word, (A, w) records,
                                    // n-leg records
    int all_orphans;                // test hook: k_rows_ctr recomputes every plane itself
    int discard;                    // k_rows_ctr drops consumed scratch lines from the L2 (UF3B_NO_DISCARD: off)
    const double *legv, *legd, *epos;   // k_centre_legs tables
    double *planes;                 // [entry][LM][NA] plane table
    double *tile3;                  // [atom][n][c][{l, m}]: neighbour-role force tiles, k_rows_nbr -> k_rows_ctr
};

template <int LM, int NA>
struct TiledShape {
    static constexpr int G = 32 / NA;                 // leg groups per contraction round
    static constexpr int NQA = (LM + 1) / 2;          // 16-byte words holding A[LM]
    static constexpr int NQ = (LM + 3 + 1) / 2;       // ... holding A[LM], w[3]
    static constexpr unsigned AWB = 16u * NQ;         // bytes of an (A, w) record
    static constexpr unsigned VNB = 80;               // bytes of an n-leg record: the four non-zero (B_n, dB_n),
                                                      // then {first basis index - n0, pad}
    static constexpr int NS = LM * (LM + 1) / 2;      // unordered pairs {l, m}
    static constexpr int PL = LM * NA;                // doubles per plane
    static constexpr int PLS = (PL + 1) & ~1;         // ... per plane of the table: whole 16-byte words (bulk copies)
    // doubles per atom of the tile buffer (3 NS per lane n), rounded up to whole 128-byte lines so that the
    // consumer can drop an atom's lines from the L2 once it has read them
    static constexpr int TS = (NA * 3 * NS + 15) / 16 * 16;
};

// Drops a 128-byte line from the L2 WITHOUT writing it back (the data becomes undefined): for scratch that
// the producer rewrites before anybody reads it again.  `p` must be 128-byte aligned.
__device__ __forceinline__ void discard_line(const void *p) {
    asm volatile("discard.global.L2 [%0], 128;" :: "l"(p) : "memory");
}
constexpr unsigned OWN_REC = 96;                      // v[4] dv[4] u[3] pad
constexpr int TL_MAX_ROW = 32;                        // longest 3-body row the kernels take
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

// compressed column of the cell ({l, m}, n) of the untrimmed window, or -1
__device__ __forceinline__ int cell_col(const BasisTab &B, const TiledGeom &tg, int l, int m, int n) {
    return __ldg(B.bin_col + tg.goff + ((tg.l0 + l) * tg.dim_m + tg.l0 + m) * tg.dim_n + tg.n0 + n);
}

// Sum of the lanes (g, n), g = 0..G-1, delivered to lane n (the other lanes get garbage).
template <int G, int NA>
__device__ __forceinline__ double fold_groups(double v) {
    double t = v;
#pragma unroll
    for (int g = 1; g < G; ++g) t += __shfl_down_sync(FULL, v, g * NA);
    return t;
}

// ---------------------------------------------------------------- k_rows_nbr
// Block = W independent warps, warp = one atom at a time.  Per-warp shared memory: the legs
// centre(e) -> a' of the atom's row, a zero word, and the chunk buffers of cg groups x ps partner
// slots (slot = position in the centre's row; the slot of the group's own atom is a dead record).
// want_f = 0 (energy row only): the planes are still needed by k_rows_ctr; they are built with the
// values-only contraction and nothing is written to the rows.
template <int LM, int NA>
__global__ void __launch_bounds__(128, 4)
k_rows_nbr(const BasisTab B, const FrameView f, const TiledGeom tg, double *__restrict__ xf, long long ld, int want_f_) {
    using S = TiledShape<LM, NA>;
    constexpr int NS = S::NS;
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int gw = blockIdx.x * nw + warp, n_gw = gridDim.x * nw;
    const bool want_f = want_f_ != 0;

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

    // ---- warp region
    unsigned char *mine = smem + tg.off_warps + (size_t)warp * (size_t)tg.warp_bytes;
    const unsigned grp_s = pin(smem_s + (unsigned)tg.off_warps + (unsigned)warp * (unsigned)tg.warp_bytes);
    const unsigned zero_s = grp_s + (unsigned)tg.off_zero;
    const unsigned aw_s = grp_s + (unsigned)tg.off_aw, vn_s = grp_s + (unsigned)tg.off_vn;
    if (lane == 0) *reinterpret_cast<double2 *>(mine + tg.off_zero) = make_double2(0.0, 0.0);
    const int ps = tg.ps, cg = tg.cg;
    const int gpp = 32 >> tg.sl_shift;                                   // groups per evaluation pass
    const int e_gg = lane >> tg.sl_shift, e_k = lane & ((1 << tg.sl_shift) - 1);
    const int c_g = lane / NA;                                           // contraction role of the lane: group,
    const int c_n = lane - c_g * NA;                                     // n
    const int c_pick = c_n < tg.na ? c_n : -(1 << 24);                   // lanes outside the window pick nothing
    const bool c_on = c_g < S::G;
    __syncwarp();

    for (int a = gw; a < f.n; a += n_gw) {
        const int row0 = __ldg(f.off3 + a);
        int n3a = __ldg(f.cnt3 + a);
        if (n3a > ps) n3a = 0;      // only behind a deferred list build whose frame will be repeated
        double fr[3][NS];           // force tile [c][{l, m}] of the lane's (g, n)
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int s = 0; s < NS; ++s) fr[c][s] = 0.0;

        // entry e of the atom's row: the row of its parent atom (the centre i), the position of a's
        // image a' in it, and that centre's leg to a'
        int my_rowi = 0, my_ni = 0, my_qa = -1;
        if (lane < n3a) {
            const size_t p = (size_t)(row0 + lane);
            const double2 tag = __ldg(reinterpret_cast<const double2 *>(tg.legd + 8 * p) + 3);
            const double2 pzt = __ldg(reinterpret_cast<const double2 *>(tg.epos + 4 * p) + 1);
            my_qa = (int)(__double_as_longlong(tag.y) >> 32);
            const long long rt = __double_as_longlong(pzt.y);
            my_rowi = (int)(rt & 0xffffffffll);
            my_ni = (int)(rt >> 32);
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
                const int rowi = __shfl_sync(FULL, my_rowi, e), ni = __shfl_sync(FULL, my_ni, e);
                const int qa = __shfl_sync(FULL, my_qa, e);
                const bool on = c_on && gi < ng && qa >= 0;
                double P[LM], Q[3][LM];
                if (want_f)
                    contract_group<LM, NA, true>(aw_s + S::AWB * (unsigned)(gi * ps), vn_s + S::VNB * (unsigned)(gi * ps),
                                                 on ? ni : 0, c_pick, zero_s, P, Q);
                else
                    contract_group<LM, NA, false>(aw_s + S::AWB * (unsigned)(gi * ps), vn_s + S::VNB * (unsigned)(gi * ps),
                                                  on ? ni : 0, c_pick, zero_s, P, Q);
                if (on) {
                    // the plane of (centre i, neighbour a') for the centre role of atom i
                    double *pl = tg.planes + (size_t)(rowi + qa) * S::PLS + c_n;
#pragma unroll
                    for (int m = 0; m < LM; ++m) pl[m * NA] = P[m];
                }
                if (on && want_f) {     // the centre's leg to a': x_a += -u dB_l (x) P + B_l (x) Q
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

        // ---- the neighbour-role force tile of the atom: lanes of different g are summed, lane n < na then
        // holds every {l, m} of its n and leaves them, 3 NS doubles in a row, in the tile buffer; k_rows_ctr
        // starts its own tile from these values and writes the rows ONCE (storing them into the rows here
        // and adding the centre role there cost k_rows_ctr a read of every 3-body column: 16 % of its stall
        // samples and 26 MB of DRAM traffic per frame)
        if (want_f) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int s = 0; s < NS; ++s) fr[c][s] = fold_groups<S::G, NA>(fr[c][s]);
            if (lane < tg.na) {
                double *dst = tg.tile3 + (size_t)a * S::TS + lane * (3 * NS);
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int s = 0; s < NS; ++s) dst[c * NS + s] = fr[c][s];
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------- k_rows_ctr
// The plane of an ORPHAN group (centre a, neighbour j whose own row does not hold a — the list
// criterion r3min < d <= r3max can differ by one ulp between the two ends of a bond, so k_rows_nbr
// never saw the group): P[m] for the lane's n, every lane evaluating the legs (j, k) itself from the
// global tables.  Rare; also the test hook tg.all_orphans.
template <int LM>
struct PlaneCol { double p[LM]; };

// (arguments by value: a reference to the kernel's parameter structs would move them to local memory)
template <int LM, int NA>
__device__ __noinline__ PlaneCol<LM> orphan_plane(const double *knots, const double *poly, int nk, double scale,
                                                  int lead, int trail, int n0, const double *epos, const double *legv,
                                                  int row0, int n3a, int j, int n) {
    PlaneCol<LM> out;
#pragma unroll
    for (int m = 0; m < LM; ++m) out.p[m] = 0.0;
    const double2 *gj = reinterpret_cast<const double2 *>(epos + 4 * (size_t)(row0 + j));
    const double2 jxy = __ldg(gj), jzt = __ldg(gj + 1);
    const Vec3 pj = {jxy.x, jxy.y, jzt.x};
    for (int k = 0; k < n3a; ++k) {
        if (k == j) continue;
        const double2 *gk = reinterpret_cast<const double2 *>(epos + 4 * (size_t)(row0 + k));
        const double2 kxy = __ldg(gk), kzt = __ldg(gk + 1);
        const Vec3 pk = {kxy.x, kxy.y, kzt.x};
        const double d = dist_rn(pj, pk);
        if (!(d >= __ldg(knots) && d <= __ldg(knots + nk - 1))) continue;        // angles.py:502-508
        double v[4], dv[4];
        const int idx = eval_leg(knots, nk, scale, poly, d, lead, trail, v, dv);
        if (idx < 0) continue;
        const int q = n - (idx - n0);
        const double vn = q == 0 ? v[0] : (q == 1 ? v[1] : (q == 2 ? v[2] : (q == 3 ? v[3] : 0.0)));
        const double *av = legv + 4 * (size_t)(row0 + k);
#pragma unroll
        for (int m = 0; m < LM; ++m) out.p[m] = fma(__ldg(av + m), vn, out.p[m]);
    }
    return out;
}

// Pair rows (bspline.evaluate_basis_functions / featurize_force_2B, bspline.py:810-895) of one atom:
// lanes evaluate 32 pairs at a time into records [q][e, fx, fy, fz] (the four non-zero basis
// functions, force parts already multiplied by 2 u: every bond is seen from both ends,
// distances.py:118-120); then lane = (slot s, q) adds entry q of the pairs s, s + 8, ... into the
// PRIVATE column accumulators of slot s — the four lanes of a pair hit four distinct columns and
// no two slots share an array, so the read-modify-writes need no ordering.  (The gather used by
// featurize.cu walks all 32 records once per column: 14 instructions per record and column block.)
constexpr int PR_SLOTS = 8;
constexpr unsigned PR_REC = 144;        // 128 bytes of entries + {first column, pad}: 16-byte words, odd multiple

struct PairTab {
    unsigned knots_s, poly_s;           // shared-window addresses of the pair's knots / pieces
    int nk, col, lead, trail;
    double scale;
};

// One pass: lane evaluates its pair (neighbour position pj, `valid`) into its record, then the
// records of the pass are added to the slots.
__device__ __forceinline__ void pair_pass(const PairTab &T, const Vec3 &pa, const Vec3 &pj, bool valid, int count,
                                          unsigned rec_s, unsigned my_slot, int lane) {
    const int s = lane >> 2, q = lane & 3;
    int col0 = -1;
    if (valid) {
        const double d = dist_rn(pa, pj);
        const double t_first = lds64(T.knots_s + 24u), t_last = lds64(T.knots_s + 8u * (unsigned)(T.nk - 4));
        if (d > t_first && d <= t_last) {       // find_interval (spline.cuh) on the staged table
            int i = 3 + (int)((d - t_first) * T.scale);
            if (i > T.nk - 5) i = T.nk - 5;
            double ti = lds64(T.knots_s + 8u * (unsigned)i);
            if (!(ti < d && d <= lds64(T.knots_s + 8u * (unsigned)i + 8u))) {
                int lo = 3, hi = T.nk - 4;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (lds64(T.knots_s + 8u * (unsigned)mid) < d) lo = mid; else hi = mid;
                }
                i = lo;
                ti = lds64(T.knots_s + 8u * (unsigned)i);
            }
            const double u = d - ti, inv2 = 2.0 / d;
            const double ux = (pj.x - pa.x) * inv2, uy = (pj.y - pa.y) * inv2, uz = (pj.z - pa.z) * inv2;
            const unsigned piece = T.poly_s + PIECE_B * (unsigned)(i - 3);
            const unsigned rec = rec_s + PR_REC * (unsigned)lane;
            const int nb = T.nk - 4;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const double2 c01 = lds128(piece + 32u * p), c23 = lds128(piece + 32u * p + 16u);
                double v = ((c23.y * u + c23.x) * u + c01.y) * u + c01.x;
                double dv = (3.0 * c23.y * u + 2.0 * c23.x) * u + c01.y;
                const int bi = i - 3 + p;
                if (bi < T.lead || bi >= nb - T.trail) { v = 0.0; dv = 0.0; }     // bspline.py:840
                sts128(rec + 32u * p, make_double2(v, dv * ux));
                sts128(rec + 32u * p + 16u, make_double2(dv * uy, dv * uz));
            }
            col0 = T.col + i - 3;
        }
    }
    asm volatile("st.shared.s32 [%0], %1;" :: "r"(rec_s + PR_REC * (unsigned)lane + 128u), "r"(col0) : "memory");
    __syncwarp();
    // consecutive pairs of a slot can touch the same column from different lanes (pair t: column c + q, pair
    // t + 8: column c' + q'): the warp-wide barrier per round orders them (uniform trip count)
    for (int t0 = 0; t0 < count; t0 += PR_SLOTS) {
        const int t = t0 + s;
        if (t < count) {
            const unsigned rec = rec_s + PR_REC * (unsigned)t;
            const int c = lds32(rec + 128u);
            if (c >= 0) {
                const unsigned ad = my_slot + 32u * (unsigned)(c + q);
                const double2 r01 = lds128(rec + 32u * q), r23 = lds128(rec + 32u * q + 16u);
                double2 a01 = lds128(ad), a23 = lds128(ad + 16u);
                a01.x += r01.x; a01.y += r01.y; a23.x += r23.x; a23.y += r23.y;
                sts128(ad, a01);
                sts128(ad + 16u, a23);
            }
        }
        __syncwarp();
    }
}

// Block = W independent warps, warp = one atom at a time.  Per-block shared memory: the pair spline
// table; per warp: PR_SLOTS private accumulator arrays [col0][e, fx, fy, fz] of the composition and
// pair columns and the records of a pair pass.
template <int LM, int NA>
__global__ void __launch_bounds__(128, 4)
k_rows_ctr(const BasisTab B, const FrameView f, const TiledGeom tg, double *__restrict__ xf, long long ld,
           double *__restrict__ partials, int want_e_, int want_f_) {
    using S = TiledShape<LM, NA>;
    constexpr int NS = S::NS;
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int gw = blockIdx.x * nw + warp, n_gw = gridDim.x * nw;
    const int F = B.n_feats;
    const bool want_e = want_e_ != 0, want_f = want_f_ != 0;

    // ---- block: pair spline table (unary basis: one pair)
    const int nk2 = __ldg(B.pair_nk), n_kn = (nk2 + 1) & ~1;
    {
        double *tab = reinterpret_cast<double *>(smem);
        const int n_po = 16 * (nk2 - 7);
        for (int k = threadIdx.x; k < nk2; k += blockDim.x) tab[k] = __ldg(B.knots2 + k);
        for (int k = threadIdx.x; k < n_po; k += blockDim.x)
            tab[n_kn + (k >> 4) * (int)(PIECE_B / 8) + (k & 15)] = __ldg(B.poly2 + k);
    }
    __syncthreads();
    const unsigned smem_s = pin(smem_addr(smem));
    PairTab pt;
    pt.knots_s = smem_s;
    pt.poly_s = smem_s + 8u * (unsigned)n_kn;
    pt.nk = nk2;
    pt.col = __ldg(B.pair_col);
    pt.lead = B.lead2;
    pt.trail = B.trail2;
    pt.scale = __ldg(B.pair_scale);

    const unsigned slot_stride = 32u * (unsigned)tg.col0;
    const unsigned slot_s = smem_s + (unsigned)tg.off_warps + (unsigned)warp * (unsigned)tg.warp_bytes;
    const unsigned rec_s = slot_s + PR_SLOTS * slot_stride;       // pair records of a pass, or (before them) the row tables:
    // row tables of the atom, staged by three TMA bulk copies (contiguous blocks of the global tables):
    // legs a -> e as values (32 B per entry) and derivatives + unit vector + tag (64 B), then the planes
    const unsigned ov_s = rec_s, od_s = rec_s + 32u * (unsigned)tg.ps, pl_s = rec_s + OWN_REC * (unsigned)tg.ps;
    const unsigned bar_s = slot_s + (unsigned)tg.warp_bytes - 16u;      // the warp's mbarrier: behind everything that is reused
    unsigned parity = 0;
    if (lane == 0) mbar_init(bar_s, 1);
    const int c_g = lane / NA, c_n = lane - c_g * NA;
    const bool c_on = c_g < S::G && c_n < tg.na;
    const double half_e = want_e ? 0.5 : 0.0;

    for (unsigned o = 16u * lane; o < PR_SLOTS * slot_stride; o += 512u) sts128(slot_s + o, make_double2(0.0, 0.0));
    double er[NS];              // energy tile {l, m} of the lane's (g, n), kept over all atoms of the warp
#pragma unroll
    for (int s = 0; s < NS; ++s) er[s] = 0.0;
    __syncwarp();

    for (int a = gw; a < f.n; a += n_gw) {
        // Every global load of the atom is issued in three batches before the arithmetic — the kernel
        // is a chain of dependent gathers otherwise (58 % of the stall samples were long-scoreboard
        // waits): (1) row bounds, (2) the lane's own 3-body entry and its pair-list entries of the
        // first two passes, (3) the planes and the pair partners' positions.
        const int row0 = __ldg(f.off3 + a);
        int n3a = __ldg(f.cnt3 + a);
        if (n3a > tg.ps) n3a = 0;   // only behind a deferred list build whose frame will be repeated
        const int r0 = __ldg(f.off2 + a), r1 = r0 + __ldg(f.cnt2 + a);
        const int sa = __ldg(f.spec + a);
        const Vec3 pa = real_position(f, a);
        // the neighbour-role tile k_rows_nbr left for this atom: the lane's 3 NS values start the lane's own
        // tile (lanes of g = 0 only: fold_groups sums over g at the end)
        double fr[3][NS];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int s = 0; s < NS; ++s) fr[c][s] = 0.0;
        if (want_f && lane < tg.na) {
            const double *src = tg.tile3 + (size_t)a * S::TS + lane * (3 * NS);
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int s = 0; s < NS; ++s) fr[c][s] = __ldcs(src + c * NS + s);
        }
        const bool pv0 = r0 + lane < r1, pv1 = r0 + 32 + lane < r1;
        const int m0 = pv0 ? __ldg(f.idx2 + r0 + lane) : 0, m1 = pv1 ? __ldg(f.idx2 + r0 + 32 + lane) : 0;
        // the row tables of the atom: bulk copies issued by lane 0 (cp.async.bulk, completion counted on the
        // warp's mbarrier), waited for behind the pair-list gathers
        if (n3a > 0 && lane == 0) {
            fence_proxy_async();        // the pair records of the previous atom lived in the same memory
            const unsigned n = (unsigned)n3a;
            mbar_expect_tx(bar_s, n * (OWN_REC + 8u * (unsigned)S::PLS));
            bulk_copy_g2s(ov_s, tg.legv + 4 * (size_t)row0, 32u * n, bar_s);
            bulk_copy_g2s(od_s, tg.legd + 8 * (size_t)row0, 64u * n, bar_s);
            bulk_copy_g2s(pl_s, tg.planes + (size_t)row0 * S::PLS, 8u * (unsigned)S::PLS * n, bar_s);
        }
        int dummy;
        Vec3 pj0 = pa, pj1 = pa;
        if (pv0) pj0 = super_position(f, m0, dummy);
        if (pv1) pj1 = super_position(f, m1, dummy);
        if (n3a > 0) {
            mbar_wait(bar_s, parity);
            parity ^= 1u;
        }
        __syncwarp();

        // ------------------------------------------------ 3-body, centre role: x_a += u_aj dB_l (x) P_j,
        // e += 1/2 B_l (x) P_j with the planes k_rows_nbr left in the table
        if (c_on && n3a > 1) {
            for (int j = c_g; j < n3a; j += S::G) {
                double2 q[6];
                q[0] = lds128(ov_s + 32u * (unsigned)j);
                q[1] = lds128(ov_s + 32u * (unsigned)j + 16u);
#pragma unroll
                for (int i = 0; i < 4; ++i) q[2 + i] = lds128(od_s + 64u * (unsigned)j + 16u * i);
                const int qa = (int)(__double_as_longlong(q[5].y) >> 32);
                double P[LM];
                if (qa >= 0 && !tg.all_orphans) {
                    const unsigned pj_ = pl_s + 8u * (unsigned)(j * S::PLS + c_n);
#pragma unroll
                    for (int m = 0; m < LM; ++m) P[m] = lds64(pj_ + 8u * (unsigned)(m * NA));
                } else {
                    const PlaneCol<LM> o = orphan_plane<LM, NA>(B.knots3 + tg.koff_n, B.poly3 + tg.poff_n, tg.nk_n, tg.scale_n,
                                                                B.lead3, B.trail3, tg.n0, tg.epos, tg.legv, row0, n3a, j, c_n);
#pragma unroll
                    for (int m = 0; m < LM; ++m) P[m] = o.p[m];
                }
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
        __syncwarp();       // the pair records reuse the memory of the row tables
        // ------------------------------------------------ 2-body and composition columns
        {
            if (lane == 0) sts64(slot_s + 32u * (unsigned)sa, lds64(slot_s + 32u * (unsigned)sa) + 1.0);   // n_el (composition.py:96-111)
            __syncwarp();
            const unsigned my_slot = slot_s + (unsigned)(lane >> 2) * slot_stride;
            if (r0 < r1) pair_pass(pt, pa, pj0, pv0, min(32, r1 - r0), rec_s, my_slot, lane);
            if (r0 + 32 < r1) pair_pass(pt, pa, pj1, pv1, min(32, r1 - r0 - 32), rec_s, my_slot, lane);
            for (int base = r0 + 64; base < r1; base += 32) {
                const bool pv = base + lane < r1;
                Vec3 pj = pa;
                if (pv) pj = super_position(f, __ldg(f.idx2 + base + lane), dummy);
                pair_pass(pt, pa, pj, pv, min(32, r1 - base), rec_s, my_slot, lane);
            }
        }
        // ------------------------------------------------ rows fx_a, fy_a, fz_a
        if (want_f) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int s = 0; s < NS; ++s) fr[c][s] = fold_groups<S::G, NA>(fr[c][s]);
            if (lane < tg.na) {         // 3-body columns: neighbour role (loaded above) + centre role, written once
#pragma unroll
                for (int l = 0; l < LM; ++l)
#pragma unroll
                    for (int m = l; m < LM; ++m) {
                        if (m < tg.la) {
                            const int col = cell_col(B, tg, l, m, lane);
                            if (col >= 0) {
                                double *dst = xf + (long long)a * ld + tg.col0 + col;
                                const long long cs = (long long)f.n * ld;
                                const int s = sym_idx(l, m, LM);
                                __stcs(dst, fr[0][s]);
                                __stcs(dst + cs, fr[1][s]);
                                __stcs(dst + 2 * cs, fr[2][s]);
                            }
                        }
                    }
            }
            // composition and pair columns: the slots' force parts are summed, stored and cleared
            for (int it = lane; it < 3 * tg.col0; it += 32) {
                const int col = it / 3, c = it - 3 * col;
                const unsigned ad = slot_s + 32u * (unsigned)col + 8u * (unsigned)(1 + c);
                double t = 0.0;
#pragma unroll
                for (int sl = 0; sl < PR_SLOTS; ++sl) {
                    t += lds64(ad + (unsigned)sl * slot_stride);
                    sts64(ad + (unsigned)sl * slot_stride, 0.0);
                }
                __stcs(xf + ((long long)c * f.n + a) * ld + col, t);     // written once: streaming
            }
        }
        // the atom's tile and planes have been consumed and are rewritten before they are read again (next
        // frame): their whole lines are dropped from the L2 instead of being written back to HBM
        if (tg.discard) {
            __syncwarp();
            if (want_f) {
                const char *t0 = reinterpret_cast<const char *>(tg.tile3 + (size_t)a * S::TS);
                for (int k = lane; k < S::TS / 16; k += 32) discard_line(t0 + 128 * k);
            }
            const size_t b0 = ((size_t)row0 * S::PLS * 8 + 127) & ~size_t(127);
            const size_t b1 = ((size_t)(row0 + n3a) * S::PLS * 8) & ~size_t(127);
            const char *p0 = reinterpret_cast<const char *>(tg.planes);
            for (size_t b = b0 + 128 * (size_t)lane; b < b1; b += 128 * 32) discard_line(p0 + b);
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
                const double t = fold_groups<S::G, NA>(er[s]);
                if (lane < tg.na && m < tg.la) {
                    const int col = cell_col(B, tg, l, m, lane);
                    if (col >= 0) mine_p[tg.col0 + col] = t;
                }
            }
        for (int col = lane; col < tg.col0; col += 32) {
            double t = 0.0;
#pragma unroll
            for (int sl = 0; sl < PR_SLOTS; ++sl) t += lds64(slot_s + (unsigned)sl * slot_stride + 32u * (unsigned)col);
            mine_p[col] = t;
        }
    }
}

// ---------------------------------------------------------------- host side
template <int LM, int NA>
static int launch_tiled(uf3b_basis *basis, const uf3b_nlist *nl, TiledGeom tg, double *x_energy, double *x_forces,
                        int64_t ld, cudaStream_t stream, int max3_in) {
    using S = TiledShape<LM, NA>;
    const int F = basis->n_feats, n = (int)nl->n;
    const bool e_dev = x_energy && is_device_pointer(x_energy);
    const bool f_dev = x_forces && is_device_pointer(x_forces);
    int dev = 0, smem_max = 0, smem_sm = 0;
    UF3B_CUDA(cudaGetDevice(&dev));
    UF3B_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    UF3B_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));

    const int max3 = std::max(max3_in, 2);
    tg.ps = max3;                   // one slot per position of the longest row
    tg.sl_shift = tg.ps <= 16 ? 4 : 5;
    tg.all_orphans = getenv("UF3B_TILED_ORPHANS") ? 1 : 0;
    tg.discard = getenv("UF3B_NO_DISCARD") ? 0 : 1;
    // a chunk of cg groups holds whole contraction rounds (G groups); the larger candidate also holds
    // whole evaluation passes (32 >> sl_shift groups) and is taken when it does not cost resident warps
    const int cg_env = getenv("UF3B_TILED_CG") ? atoi(getenv("UF3B_TILED_CG")) : 0;
    const int w_env = getenv("UF3B_TILED_WARPS") ? atoi(getenv("UF3B_TILED_WARPS")) : 0;
    const int gpp = 32 >> tg.sl_shift;
    const int rows = (max3 + 1) & ~1;
    const size_t tab_bytes = 8 * (size_t)((tg.nk_n + 1) & ~1) + (size_t)PIECE_B * (tg.nk_n - 7);
    tg.off_warps = (int)((tab_bytes + 15) & ~size_t(15));
    tg.off_zero = rows * (int)OWN_REC;
    tg.off_aw = tg.off_zero + 16;
    auto layout = [&](int cg, int &warps, int &blocks) {
        tg.cg = cg;
        // two records of readable padding behind each array (the contraction loads ahead)
        const size_t aw_bytes = ((size_t)cg * tg.ps + 2) * S::AWB;
        tg.off_vn = tg.off_aw + (int)aw_bytes;
        const size_t chunk_bytes = aw_bytes + ((size_t)cg * tg.ps + 2) * S::VNB;
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
    auto k_nbr = k_rows_nbr<LM, NA>;
    auto k_ctr = k_rows_ctr<LM, NA>;
    const size_t smem_n = (size_t)tg.off_warps + (size_t)warps * tg.warp_bytes;
    // k_rows_ctr: [pair spline table][per warp: PR_SLOTS x col0 accumulator quads, 32 pair records]
    const int warps_c = 4;
    const int nk2 = basis->h_pair_nk0;
    const size_t tab_c = ((8 * (size_t)((nk2 + 1) & ~1) + (size_t)PIECE_B * (nk2 - 7)) + 15) & ~size_t(15);
    const size_t rows_c = (size_t)OWN_REC * tg.ps + 8 * (size_t)S::PLS * tg.ps;
    // [slot accumulators][pair records of a pass | row tables of the atom][mbarrier of the bulk copies]
    const size_t warp_c = (((size_t)PR_SLOTS * 32 * tg.col0 + std::max<size_t>(32 * PR_REC, rows_c) + 15) & ~size_t(15)) + 16;
    const size_t smem_c = tab_c + (size_t)warps_c * warp_c;
    if (smem_c > (size_t)smem_max) return 1;
    TiledGeom tgc = tg;
    tgc.off_warps = (int)tab_c;
    tgc.warp_bytes = (int)warp_c;
    UF3B_CUDA(ensure_dynamic_smem((const void *)k_nbr, smem_n));
    UF3B_CUDA(ensure_dynamic_smem((const void *)k_ctr, smem_c));
    int per_sm_n = 1, per_sm_c = 1;
    UF3B_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_n, k_nbr, warps * 32, smem_n));
    UF3B_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_c, k_ctr, warps_c * 32, smem_c));
    const int share = std::max(1, basis->frames_in_flight);
    int grid_n = std::max(1, sm_count() * std::max(per_sm_n, 1) / share);
    grid_n = std::min(grid_n, (n + warps - 1) / warps);
    int grid_c = std::max(1, sm_count() * std::max(per_sm_c, 1) / share);
    grid_c = std::min(grid_c, (n + warps_c - 1) / warps_c);
    const int n_gw = grid_c * warps_c;

    // tables of k_centre_legs and the planes: one slot per list entry (rows live in claim regions of idx3)
    const size_t entries = nl->idx3.cap;
    UF3B_CUDA(basis->legv.reserve(4 * entries));
    UF3B_CUDA(basis->legd.reserve(8 * entries));
    UF3B_CUDA(basis->epos.reserve(4 * entries));
    UF3B_CUDA(basis->planes.reserve((size_t)S::PLS * entries));
    if (x_forces) UF3B_CUDA(basis->tile3.reserve((size_t)n * S::TS + 16));
    tg.legv = basis->legv.p;
    tg.legd = basis->legd.p;
    tg.epos = basis->epos.p;
    tg.planes = basis->planes.p;
    tg.tile3 = basis->tile3.p;

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
    UF3B_LAUNCH(k_nbr, grid_n, warps * 32, smem_n, stream, basis->tab, view, tg, d_xf, d_ld, x_forces ? 1 : 0);
    tgc.legv = tg.legv; tgc.legd = tg.legd; tgc.epos = tg.epos; tgc.planes = tg.planes; tgc.tile3 = tg.tile3;
    UF3B_LAUNCH(k_ctr, grid_c, warps_c * 32, smem_c, stream, basis->tab, view, tgc, d_xf, d_ld, basis->partials.p,
                x_energy ? 1 : 0, x_forces ? 1 : 0);
    if (g_timing) UF3B_CUDA(cudaEventRecord(ev1, stream));
    if (x_energy)
        if (int rc = launch_energy_row(basis->partials.p, n_gw, F, d_xe, stream)) return rc;
    return finish_featurize(basis, x_energy, x_forces, ld, d_xe, d_xf, F, n, e_dev, f_dev, stream, ev0, ev1);
}

// Takes the frame if the basis fits the tiled kernels: unary trio of symmetry 2 with unit folding
// weights, untrimmed grid la x la x na with la <= 4, na <= 10, rows of at most 32 entries.
// Returns 1 when it does not apply (the caller goes on to the other paths); error codes are <= 0.
int featurize_tiled(uf3b_basis *basis, const uf3b_nlist *nl, double *x_energy, double *x_forces, int64_t ld,
                    cudaStream_t stream, bool deferred) {
    const BasisTab &T = basis->tab;
    if (T.n_trios != 1 || T.ne != 1 || !T.unit_weights || basis->no_tile || getenv("UF3B_NO_TILED") || getenv("UF3B_NO_LEGS")
        || getenv("UF3B_PLANES"))
        return 1;
    // deferred list build: the longest row is not known on the host yet; the kernels are sized by the
    // previous frame's and skip longer rows (the device flag of the build then marks the frame invalid)
    const int max3_in = deferred ? nl->max3_hint : nl->max3;
    if (basis->h_trio_sym[0] != 2 || max3_in > TL_MAX_ROW) return 1;
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
    if (deferred) {
        const_cast<uf3b_nlist *>(nl)->hint_used = true;
        const_cast<uf3b_nlist *>(nl)->consumed_pending = true;
    }
    if (tg.la <= 2 && tg.na <= 7) return launch_tiled<2, 7>(basis, nl, tg, x_energy, x_forces, ld, stream, max3_in);
    if (tg.la <= 3 && tg.na <= 9) return launch_tiled<3, 9>(basis, nl, tg, x_energy, x_forces, ld, stream, max3_in);
    return launch_tiled<4, 10>(basis, nl, tg, x_energy, x_forces, ld, stream, max3_in);
}

}  // namespace uf3b
