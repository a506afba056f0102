// featurize_multi.cu — fit path, leg-grouped kernel for ANY chemical system: several species,
// trios of symmetry 1 or 2, 3-body rows of any length.
//
// The reference buckets the triangles of a centre per trio interaction after ordering the two
// neighbours by atomic number (representation/angles.py:460-496) and scatters every triangle's 4x4x4
// block of basis products into dense per-atom grids (angles.py:104-139, :235-286).  Here the rows of
// an atom factor per LEG GROUP and PARTNER CLASS, as in featurize_tiled.cu but without its
// restrictions (one species, symmetry 2, rows of at most 32 entries):
//
//   group  = (centre c, one of its neighbours g): centre role   c = a,  g = j   for every entry j of a's row
//                                                 neighbour role c = i,  g = a'  for every centre i in a's row
//   class  = the partners k of the group that fall into the same trio interaction with the same leg
//            order: species of k, and — only for a same-species trio of symmetry 1 — whether k lies
//            behind or before g in the centre's row (angles.py:474-488: legs l, m by atomic number,
//            ties by supercell index)
//   plane  P[x, n]   = sum_k B_x(r_ck) B_n(r_gk)                    x = basis index of the partner's leg
//          Q_c[x, n] = sum_k w_gk,c B_x(r_ck) B'_n(r_gk)            (neighbour role; w = unit vector g -> k)
//   rows   centre role     x_a += u_cg B'_y(r_cg) (x) P,   e += 1/2 B_y(r_cg) (x) P
//          neighbour role  x_a += -u_cg B'_y(r_cg) (x) P + B_y(r_cg) (x) Q
//          at bin (l, m, n) = (y, x, n) when g is the trio's l leg, (x, y, n) when it is the m leg.
//
// Warp = one atom at a time, accumulators [column][e, fx, fy, fz] in the warp's shared memory (no
// atomics, bit-reproducible).  Lanes evaluate the legs of 32 partners at a time into sparse records
// (4 non-zero values + first basis index); then lane = cell (x, n) of the plane (up to 4 cells per
// lane) contracts the records with P, Qx, Qy, Qz in registers, and adds the four non-zero y terms of
// the group's own leg to the compressed columns of its cells (bin_col).  Every triangle is evaluated in
// the frame of its REAL centre, so list order, leg order and every distance are the reference's
// real-centre enumeration (triangle.cuh).
#include <algorithm>
#include <cstdlib>

#include "featurize_common.cuh"

namespace uf3b {

constexpr unsigned MR_A = 48;         // partner-leg record: v[4], {first index - first kept index, pad}
constexpr unsigned MR_B = 80;         // n-leg record: (v, dv)[4], {first index - first kept index, pad}
constexpr unsigned MR_W = 32;         // unit vector g -> k, pad
constexpr unsigned MR_REC = MR_A + MR_B + MR_W;

struct MultiGeom {
    int warp_bytes;         // per-warp shared memory: [acc 32 F][zero quad][32 records]
    int off_zero, off_rec;
};

// One leg of a trio: knots, pieces and the kept (untrimmed) index range.
struct LegTab {
    const double *knots, *poly;
    int nk, first, count;   // kept basis indices [first, first + count)
    double scale;
};

__device__ __forceinline__ LegTab leg_tab(const BasisTab &B, int trio, int leg) {
    LegTab T;
    const int s = 3 * trio + leg;
    T.nk = __ldg(B.trio_nk + s);
    T.knots = B.knots3 + __ldg(B.trio_koff + s);
    T.poly = B.poly3 + __ldg(B.trio_poff + s);
    T.scale = __ldg(B.trio_scale + s);
    T.first = B.lead3;
    T.count = max(0, T.nk - 4 - B.lead3 - B.trail3);
    return T;
}

// inclusive leg filter of the reference (angles.py:502-508), then the four non-zero basis functions
// with the trims applied (angles.py:554-565); -1: the triangle contributes nothing
__device__ __forceinline__ int eval_leg_of(const BasisTab &B, const LegTab &T, double d, double (&v)[4], double (&dv)[4]) {
    if (!(d >= __ldg(T.knots) && d <= __ldg(T.knots + T.nk - 1))) return -1;
    return eval_leg(T.knots, T.nk, T.scale, T.poly, d, B.lead3, B.trail3, v, dv);
}

template <int NCB>
__global__ void __launch_bounds__(128, 2)
k_rows_multi(const BasisTab B, const FrameView f, const MultiGeom mg, double *__restrict__ xf, long long ld,
             double *__restrict__ partials, int want_e_, int want_f_) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int gw = blockIdx.x * nw + warp, n_gw = gridDim.x * nw;
    const int F = B.n_feats;
    const bool want_e = want_e_ != 0, want_f = want_f_ != 0;
    unsigned char *mine = smem + (size_t)warp * (size_t)mg.warp_bytes;
    double *acc = reinterpret_cast<double *>(mine);
    const unsigned mine_s = pin(smem_addr(mine));
    const unsigned zero_s = mine_s + (unsigned)mg.off_zero, rec_s = mine_s + (unsigned)mg.off_rec;
    PairRec *prec = reinterpret_cast<PairRec *>(mine + mg.off_rec);
    const double half_e = want_e ? 0.5 : 0.0;

    for (int k = lane; k < 4 * F; k += 32) acc[k] = 0.0;
    if (lane == 0) *reinterpret_cast<double2 *>(mine + mg.off_zero) = make_double2(0.0, 0.0);
    __syncwarp();

    for (int a = gw; a < f.n; a += n_gw) {
        const int sa = __ldg(f.spec + a);
        const Vec3 pa = real_position(f, a);
        if (lane == 0) acc[4 * sa] += 1.0;      // composition column n_el (composition.py:96-111)
        __syncwarp();
        two_body_rows(B, f, a, sa, pa, acc, prec, lane);

        const int row0 = __ldg(f.off3 + a), n3a = B.n_trios > 0 ? __ldg(f.cnt3 + a) : 0;
        for (int role = 0; role < (want_f ? 2 : 1); ++role) {           // 0: a is the centre, 1: a is a neighbour
            if (role == 0 && n3a < 2) continue;
            for (int e = 0; e < n3a; ++e) {
                // ---- the group: centre c (real atom), its row, the neighbour g at position qg of that row
                int rowc = row0, nc = n3a, qg = e, sc = sa;
                Vec3 pc = pa;
                int mgi = __ldg(f.idx3 + row0 + e);
                if (role == 1) {
                    const int gimg = image_of(f, mgi);
                    const int ci = mgi - gimg * f.n;
                    const int apr = __ldg(f.img_inv + gimg) * f.n + a;       // a as the centre sees it
                    rowc = __ldg(f.off3 + ci);
                    nc = __ldg(f.cnt3 + ci);
                    sc = __ldg(f.spec + ci);
                    pc = real_position(f, ci);
                    qg = -1;
                    for (int k0 = 0; k0 < nc && qg < 0; k0 += 32) {
                        const int its = k0 + lane < nc ? __ldg(f.idx3 + rowc + k0 + lane) : -1;
                        const unsigned hit = __ballot_sync(FULL, its == apr);
                        if (hit) qg = k0 + __ffs(hit) - 1;
                    }
                    if (qg < 0) continue;           // one-ulp asymmetry of the list criterion: i does not list a'
                    mgi = apr;
                }
                int ag;
                const Vec3 pg = super_position(f, mgi, ag);
                const int sg = __ldg(f.spec + ag);
                const double dcg = dist_rn(pc, pg);
                const double icg = fast_rcp(dcg);
                const double ug[3] = {(pg.x - pc.x) * icg, (pg.y - pc.y) * icg, (pg.z - pc.z) * icg};

                // partners of the first pass: position, species and the two distances once for all classes
                int sk0 = -1;
                Vec3 pk0 = pc;
                double dck0 = 0.0, dgk0 = 0.0;
                if (lane < nc && lane != qg) {
                    int ak;
                    pk0 = super_position(f, __ldg(f.idx3 + rowc + lane), ak);
                    sk0 = __ldg(f.spec + ak);
                    dck0 = dist_rn(pc, pk0);
                    dgk0 = dist_rn(pg, pk0);
                }

                for (int sk = 0; sk < B.ne; ++sk) {
                    const int trio = sc * B.n_pairs + pair_index(B.ne, sg, sk);
                    const int n_sub = (sk == sg && __ldg(B.trio_sym + trio) == 1) ? 2 : 1;
                    for (int sub = 0; sub < n_sub; ++sub) {
                        const bool g_is_l = sg < sk || (sg == sk && sub == 0);
                        const LegTab tg_ = leg_tab(B, trio, g_is_l ? 0 : 1);     // the group's own leg (c, g)
                        const LegTab tk = leg_tab(B, trio, g_is_l ? 1 : 0);      // the partner's leg (c, k)
                        const LegTab tn = leg_tab(B, trio, 2);                   // the leg (g, k)
                        double gv[4], gd[4];
                        const int gi = eval_leg_of(B, tg_, dcg, gv, gd);
                        if (gi < 0 || tk.count < 1 || tn.count < 1) continue;
                        const int n_cells = tk.count * tn.count;
                        int cx[NCB], cn[NCB];               // the lane's cells (x, n), relative to the kept ranges
#pragma unroll
                        for (int b = 0; b < NCB; ++b) {
                            const int cell = lane + 32 * b;
                            const bool ok = cell < n_cells;
                            cx[b] = ok ? cell / tn.count : -(1 << 24);
                            cn[b] = ok ? cell - cx[b] * tn.count : -(1 << 24);
                        }
                        double P[NCB], Q[3][NCB];
#pragma unroll
                        for (int b = 0; b < NCB; ++b) P[b] = Q[0][b] = Q[1][b] = Q[2][b] = 0.0;

                        for (int k0 = 0; k0 < nc; k0 += 32) {
                            // ---- evaluation pass: lane = partner k0 + lane, valid records compacted
                            const int k = k0 + lane;
                            int sk_ = sk0;
                            Vec3 pk = pk0;
                            double dck = dck0, dgk = dgk0;
                            if (k0 > 0) {
                                sk_ = -1;
                                if (k < nc && k != qg) {
                                    int ak;
                                    pk = super_position(f, __ldg(f.idx3 + rowc + k), ak);
                                    sk_ = __ldg(f.spec + ak);
                                    dck = dist_rn(pc, pk);
                                    dgk = dist_rn(pg, pk);
                                }
                            }
                            bool ok = sk_ == sk && (n_sub == 1 || (sub == 0 ? k > qg : k < qg));
                            double av[4], ad[4], bv[4], bd[4];
                            int ia = -1, ib = -1;
                            if (ok) {
                                ia = eval_leg_of(B, tk, dck, av, ad);
                                ib = ia < 0 ? -1 : eval_leg_of(B, tn, dgk, bv, bd);
                                ok = ia >= 0 && ib >= 0;
                            }
                            const unsigned live = __ballot_sync(FULL, ok);
                            if (ok) {
                                const unsigned slot = (unsigned)__popc(live & ((1u << lane) - 1u));
                                const unsigned r = rec_s + MR_REC * slot;
                                sts128(r, make_double2(av[0], av[1]));
                                sts128(r + 16, make_double2(av[2], av[3]));
                                asm volatile("st.shared.s32 [%0], %1;" :: "r"(r + 32u), "r"(ia - tk.first) : "memory");
#pragma unroll
                                for (int p = 0; p < 4; ++p) sts128(r + MR_A + 16u * p, make_double2(bv[p], bd[p]));
                                asm volatile("st.shared.s32 [%0], %1;" :: "r"(r + MR_A + 64u), "r"(ib - tn.first) : "memory");
                                const double inv = fast_rcp(dgk);
                                sts128(r + MR_A + MR_B, make_double2((pk.x - pg.x) * inv, (pk.y - pg.y) * inv));
                                sts64(r + MR_A + MR_B + 16, (pk.z - pg.z) * inv);
                            }
                            __syncwarp();
                            // ---- contraction: every lane adds the live records to its cells
                            const int n_live = __popc(live);
                            for (int s = 0; s < n_live; ++s) {
                                const unsigned r = rec_s + MR_REC * (unsigned)s;
                                const int ra = lds32(r + 32u), rb = lds32(r + MR_A + 64u);
                                if (role == 1) {
                                    const double2 w01 = lds128(r + MR_A + MR_B);
                                    const double w2 = lds64(r + MR_A + MR_B + 16);
#pragma unroll
                                    for (int b = 0; b < NCB; ++b) {
                                        const unsigned qa = (unsigned)(cx[b] - ra), qb = (unsigned)(cn[b] - rb);
                                        const double x = lds64(qa < 4u ? r + 8u * qa : zero_s);
                                        const double2 y = lds128(qb < 4u ? r + MR_A + 16u * qb : zero_s);
                                        const double t = x * y.y;
                                        P[b] = fma(x, y.x, P[b]);
                                        Q[0][b] = fma(w01.x, t, Q[0][b]);
                                        Q[1][b] = fma(w01.y, t, Q[1][b]);
                                        Q[2][b] = fma(w2, t, Q[2][b]);
                                    }
                                } else {
#pragma unroll
                                    for (int b = 0; b < NCB; ++b) {
                                        const unsigned qa = (unsigned)(cx[b] - ra), qb = (unsigned)(cn[b] - rb);
                                        const double x = lds64(qa < 4u ? r + 8u * qa : zero_s);
                                        const double y = lds64(qb < 4u ? r + MR_A + 16u * qb : zero_s);
                                        P[b] = fma(x, y, P[b]);
                                    }
                                }
                            }
                            __syncwarp();
                        }

                        // ---- the group's own leg (x) planes -> compressed columns of the lane's cells
                        const int goff = __ldg(B.trio_goff + trio), col0 = __ldg(B.trio_col + trio);
                        const int dim_m = __ldg(B.trio_nk + 3 * trio + 1) - 4, dim_n = tn.nk - 4;
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            const int y = gi + p;           // basis index of the group's leg; trimmed ones are zero
                            if (y < tg_.first || y >= tg_.first + tg_.count) continue;
                            const double hv = half_e * gv[p];
                            const double d0 = ug[0] * gd[p], d1 = ug[1] * gd[p], d2 = ug[2] * gd[p];
#pragma unroll
                            for (int b = 0; b < NCB; ++b) {
                                if (cx[b] < 0) continue;
                                const int x = tk.first + cx[b], n = tn.first + cn[b];
                                const int bin = goff + ((g_is_l ? y : x) * dim_m + (g_is_l ? x : y)) * dim_n + n;
                                const int col = __ldg(B.bin_col + bin);
                                if (col < 0) continue;
                                const double w = B.unit_weights ? 1.0 : __ldg(B.bin_w + bin);
                                double2 *dst = reinterpret_cast<double2 *>(acc + 4 * (size_t)(col0 + col));
                                double2 lo = dst[0], hi = dst[1];
                                if (role == 0) {
                                    lo.x = fma(w * hv, P[b], lo.x);
                                    lo.y = fma(w * d0, P[b], lo.y);
                                    hi.x = fma(w * d1, P[b], hi.x);
                                    hi.y = fma(w * d2, P[b], hi.y);
                                } else {
                                    lo.y += w * (gv[p] * Q[0][b] - d0 * P[b]);
                                    hi.x += w * (gv[p] * Q[1][b] - d1 * P[b]);
                                    hi.y += w * (gv[p] * Q[2][b] - d2 * P[b]);
                                }
                                dst[0] = lo;
                                dst[1] = hi;
                            }
                            __syncwarp();       // two y values of a lane can fold onto one column (symmetry 2)
                        }
                    }
                }
            }
        }

        // ---- rows fx_a, fy_a, fz_a
        __syncwarp();
        if (want_f) {
            for (int col = lane; col < F; col += 32) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    __stcs(xf + ((long long)c * f.n + a) * ld + col, acc[4 * col + 1 + c]);      // written once: streaming
                    acc[4 * col + 1 + c] = 0.0;
                }
            }
        }
        __syncwarp();
    }
    if (want_e)
        for (int col = lane; col < F; col += 32) partials[(size_t)gw * F + col] = acc[4 * col];
}

// ---------------------------------------------------------------- k_rows_multi2
// The same factorisation with the loops turned inside out for small planes (<= 32 cells, kept extent of
// the l / m legs <= LA): KIND = (centre species, group species, partner species, leg order) is the outer
// loop and the groups of that kind the inner one, so that
//   * everything that depends on the kind only — the three spline tables, the lane's cell and its LA
//     compressed columns (bin_col) — is set up once per kind instead of once per group and class, and
//   * the outer products with the group's own leg are accumulated in REGISTERS, R[y][e, fx, fy, fz] for
//     the lane's cell, over all groups of the kind, and folded into the shared-memory columns once per
//     kind (k_rows_multi folds after every group: 16 read-modify-writes per lane and group).
// A per-warp table of the atom's row (ghost position, species, and for the neighbour role the centre's
// row and the position of a' in it, found once by binary search) replaces the per-group gathers.
constexpr unsigned ME_REC = 48;       // row entry: x y z | {species, parent atom} | {row start, row length} | {qg, image of a'}
constexpr int ME_MAX = 64;            // rows the entry table holds; longer rows take k_rows_multi

struct Multi2Geom {
    int warp_bytes;
    int off_zero, off_ent, off_rec;
    int col3;               // first 3-body column
    int n_knots3, n_poly3;  // doubles of the 3-body spline tables staged in front of the warps' regions
    int off_warps;          // bytes of that block-shared table
};
constexpr int M2_PS = 18;   // doubles between the polynomial pieces of the staged table (spline.cuh: bank spread)

// leg_tab on the tables staged in shared memory
__device__ __forceinline__ LegTab leg_tab_s(const BasisTab &B, const double *knots_s, const double *poly_s, int trio, int leg) {
    LegTab T;
    const int s = 3 * trio + leg;
    T.nk = __ldg(B.trio_nk + s);
    T.knots = knots_s + __ldg(B.trio_koff + s);
    T.poly = poly_s + __ldg(B.trio_poff + s) / 16 * M2_PS;
    T.scale = __ldg(B.trio_scale + s);
    T.first = B.lead3;
    T.count = max(0, T.nk - 4 - B.lead3 - B.trail3);
    return T;
}
__device__ __forceinline__ int eval_leg_s(const BasisTab &B, const LegTab &T, double d, double (&v)[4], double (&dv)[4]) {
    if (!(d >= T.knots[0] && d <= T.knots[T.nk - 1])) return -1;          // angles.py:502-508
    return eval_leg<false, M2_PS>(T.knots, T.nk, T.scale, T.poly, d, B.lead3, B.trail3, v, dv);
}
#ifndef MULTI2_BLOCKS
#define MULTI2_BLOCKS 3     // resident blocks of 4 warps the register budget is cut for (170 registers)
#endif

template <int LA, int ROLE>
__device__ __forceinline__ void add_own_leg(double (&R)[LA][4], int rel, const double (&gv)[4], const double (&gd)[4],
                                            const double (&ug)[3], double P, const double (&Q)[3], double half_e) {
    // rel = first basis index of the group's leg - first kept index; entries outside [0, LA) are trimmed (zero)
#define UF3B_OWN(REL)                                                                         \
    case REL:                                                                                 \
        _Pragma("unroll") for (int p = 0; p < 4; ++p) {                                       \
            constexpr int base = REL;                                                         \
            const int y = base + p;                                                           \
            if (y >= 0 && y < LA) {                                                           \
                const double dP = gd[p] * P;                                                  \
                if (ROLE == 0) {                                                              \
                    R[y < 0 ? 0 : (y < LA ? y : 0)][0] = fma(half_e * gv[p], P, R[y < 0 ? 0 : (y < LA ? y : 0)][0]); \
                    R[y < 0 ? 0 : (y < LA ? y : 0)][1] = fma(ug[0], dP, R[y < 0 ? 0 : (y < LA ? y : 0)][1]);         \
                    R[y < 0 ? 0 : (y < LA ? y : 0)][2] = fma(ug[1], dP, R[y < 0 ? 0 : (y < LA ? y : 0)][2]);         \
                    R[y < 0 ? 0 : (y < LA ? y : 0)][3] = fma(ug[2], dP, R[y < 0 ? 0 : (y < LA ? y : 0)][3]);         \
                } else {                                                                      \
                    R[y < 0 ? 0 : (y < LA ? y : 0)][1] += gv[p] * Q[0] - ug[0] * dP;          \
                    R[y < 0 ? 0 : (y < LA ? y : 0)][2] += gv[p] * Q[1] - ug[1] * dP;          \
                    R[y < 0 ? 0 : (y < LA ? y : 0)][3] += gv[p] * Q[2] - ug[2] * dP;          \
                }                                                                             \
            }                                                                                 \
        }                                                                                     \
        break;
    switch (rel) {
        UF3B_OWN(-3) UF3B_OWN(-2) UF3B_OWN(-1) UF3B_OWN(0) UF3B_OWN(1) UF3B_OWN(2) UF3B_OWN(3)
        UF3B_OWN(4) UF3B_OWN(5) UF3B_OWN(6) UF3B_OWN(7)
        default: break;
    }
#undef UF3B_OWN
}

template <int LA>
__global__ void __launch_bounds__(128, MULTI2_BLOCKS)
k_rows_multi2(const BasisTab B, const FrameView f, const Multi2Geom mg, double *__restrict__ xf, long long ld,
              double *__restrict__ partials, double *__restrict__ gacc, int want_e_, int want_f_) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int gw = blockIdx.x * nw + warp, n_gw = gridDim.x * nw;
    const int F = B.n_feats;
    const bool want_e = want_e_ != 0, want_f = want_f_ != 0;
    // ---- block: the 3-body spline tables (knots, then pieces M2_PS doubles apart)
    const double *knots_s = reinterpret_cast<const double *>(smem);
    const double *poly_s = knots_s + ((mg.n_knots3 + 1) & ~1);
    {
        double *kd = reinterpret_cast<double *>(smem), *pd = kd + ((mg.n_knots3 + 1) & ~1);
        for (int i = threadIdx.x; i < mg.n_knots3; i += blockDim.x) kd[i] = __ldg(B.knots3 + i);
        for (int i = threadIdx.x; i < mg.n_poly3; i += blockDim.x) pd[(i >> 4) * M2_PS + (i & 15)] = __ldg(B.poly3 + i);
    }
    __syncthreads();
    unsigned char *mine = smem + mg.off_warps + (size_t)warp * (size_t)mg.warp_bytes;
    // composition and pair columns [0, col3) accumulate in shared memory (every pair touches them); the 3-body
    // columns are only touched when a kind's registers are folded, a few times per atom: they live in a
    // per-warp global scratch row (L2-resident), which keeps the shared memory of a warp under 10 KB
    double *acc = reinterpret_cast<double *>(mine);
    const int col3 = mg.col3;
    double *acc3 = gacc + (size_t)gw * 4 * (size_t)F;
    const unsigned mine_s = pin(smem_addr(mine));
    const unsigned zero_s = mine_s + (unsigned)mg.off_zero, ent_s = mine_s + (unsigned)mg.off_ent;
    const unsigned rec_s = mine_s + (unsigned)mg.off_rec;
    PairRec *prec = reinterpret_cast<PairRec *>(mine + mg.off_rec);
    const double half_e = want_e ? 0.5 : 0.0;

    for (int k = lane; k < 4 * col3; k += 32) acc[k] = 0.0;
    for (int k = 4 * col3 + lane; k < 4 * F; k += 32) acc3[k] = 0.0;
    if (lane == 0) *reinterpret_cast<double2 *>(mine + mg.off_zero) = make_double2(0.0, 0.0);
    __syncwarp();

    for (int a = gw; a < f.n; a += n_gw) {
        const int sa = __ldg(f.spec + a);
        const Vec3 pa = real_position(f, a);
        if (lane == 0) acc[4 * sa] += 1.0;      // composition column n_el (composition.py:96-111)
        __syncwarp();
        two_body_rows(B, f, a, sa, pa, acc, prec, lane);

        const int row0 = __ldg(f.off3 + a);
        int n3a = B.n_trios > 0 ? __ldg(f.cnt3 + a) : 0;
        if (n3a > ME_MAX) n3a = 0;              // the host sends frames with longer rows to k_rows_multi
        // ---- table of the atom's row
        for (int e = lane; e < n3a; e += 32) {
            const int m = __ldg(f.idx3 + row0 + e);
            const int gimg = image_of(f, m);
            const int ci = m - gimg * f.n;
            const int ginv = __ldg(f.img_inv + gimg);
            const int apr = ginv * f.n + a;                               // a as the entry's parent atom sees it
            const int rowi = __ldg(f.off3 + ci), ni = __ldg(f.cnt3 + ci);
            int lo = 0, hi = ni;                                          // rows are sorted by supercell index
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(f.idx3 + rowi + mid) < apr) lo = mid + 1; else hi = mid;
            }
            const int qa = (lo < ni && __ldg(f.idx3 + rowi + lo) == apr) ? lo : -1;   // -1: one-ulp asymmetry of the list
            int dummy;
            const Vec3 pm = super_position(f, m, dummy);
            const unsigned r = ent_s + ME_REC * (unsigned)e;
            sts128(r, make_double2(pm.x, pm.y));
            sts128(r + 16, make_double2(pm.z, __longlong_as_double(((long long)__ldg(f.spec + ci) << 32) | (unsigned)ci)));
            sts128(r + 32, make_double2(__longlong_as_double(((long long)ni << 32) | (unsigned)rowi),
                                        __longlong_as_double(((long long)ginv << 32) | (unsigned)qa)));
        }
        __syncwarp();

        for (int role = 0; role < (want_f ? 2 : 1); ++role) {           // 0: a is the centre, 1: a is a neighbour
            if (role == 0 && n3a < 2) continue;
            // kinds: role 0 (centre a):     sc = sa, sg = species of the entry, any sk
            //        role 1 (neighbour a'): sc = species of the entry's parent atom, sg = sa, any sk
            for (int sx = 0; sx < B.ne; ++sx) {
                const int sc = role == 0 ? sa : sx, sg = role == 0 ? sx : sa;
                for (int sk = 0; sk < B.ne; ++sk) {
                    const int trio = sc * B.n_pairs + pair_index(B.ne, sg, sk);
                    const int n_sub = (sk == sg && __ldg(B.trio_sym + trio) == 1) ? 2 : 1;
                    for (int sub = 0; sub < n_sub; ++sub) {
                        const bool g_is_l = sg < sk || (sg == sk && sub == 0);
                        const LegTab tg_ = leg_tab_s(B, knots_s, poly_s, trio, g_is_l ? 0 : 1);     // the group's own leg (c, g)
                        const LegTab tk = leg_tab_s(B, knots_s, poly_s, trio, g_is_l ? 1 : 0);      // the partner's leg (c, k)
                        const LegTab tn = leg_tab_s(B, knots_s, poly_s, trio, 2);                   // the leg (g, k)
                        if (tg_.count < 1 || tk.count < 1 || tn.count < 1) continue;
                        const int n_cells = tk.count * tn.count;
                        const bool has_cell = lane < n_cells;
                        const int cx = has_cell ? lane / tn.count : -(1 << 24);
                        const int cn = has_cell ? lane - cx * tn.count : -(1 << 24);
                        double R[LA][4];
#pragma unroll
                        for (int y = 0; y < LA; ++y) R[y][0] = R[y][1] = R[y][2] = R[y][3] = 0.0;
                        bool any = false;

                        for (int e = 0; e < n3a; ++e) {
                            const unsigned er = ent_s + ME_REC * (unsigned)e;
                            const double2 e1 = lds128(er + 16);
                            const long long tag = __double_as_longlong(e1.y);
                            if ((int)(tag >> 32) != sx) continue;                 // species of the entry's parent atom
                            // ---- the group: centre c, its row, the neighbour g at position qg of that row
                            int rowc = row0, nc = n3a, qg = e;
                            Vec3 pc = pa, pg;
                            if (role == 0) {
                                const double2 e0 = lds128(er);
                                pg.x = e0.x; pg.y = e0.y; pg.z = e1.x;
                            } else {
                                const double2 e2 = lds128(er + 32);
                                const long long rt = __double_as_longlong(e2.x), qt = __double_as_longlong(e2.y);
                                qg = (int)(qt & 0xffffffffll);
                                if (qg < 0) continue;       // the centre does not list a' (one-ulp asymmetry)
                                rowc = (int)(rt & 0xffffffffll);
                                nc = (int)(rt >> 32);
                                const int ci = (int)(tag & 0xffffffffll), ginv = (int)(qt >> 32);
                                pc = real_position(f, ci);
                                pg.x = __dadd_rn(pa.x, __ldg(f.img_off + 3 * ginv + 0));      // data/geometry.py:146-147
                                pg.y = __dadd_rn(pa.y, __ldg(f.img_off + 3 * ginv + 1));
                                pg.z = __dadd_rn(pa.z, __ldg(f.img_off + 3 * ginv + 2));
                            }
                            const double dcg = dist_rn(pc, pg);
                            double gv[4], gd[4];
                            const int gi = eval_leg_s(B, tg_, dcg, gv, gd);
                            if (gi < 0) continue;
                            const double icg = fast_rcp(dcg);
                            const double ug[3] = {(pg.x - pc.x) * icg, (pg.y - pc.y) * icg, (pg.z - pc.z) * icg};
                            double P = 0.0, Q[3] = {0.0, 0.0, 0.0};
                            for (int k0 = 0; k0 < nc; k0 += 32) {
                                // ---- evaluation pass: lane = partner k0 + lane, live records compacted
                                const int k = k0 + lane;
                                bool ok = k < nc && k != qg && (n_sub == 1 || (sub == 0 ? k > qg : k < qg));
                                Vec3 pk = pc;
                                if (ok) {
                                    if (role == 0) {
                                        const unsigned kr = ent_s + ME_REC * (unsigned)k;
                                        const double2 k0_ = lds128(kr), k1_ = lds128(kr + 16);
                                        pk.x = k0_.x; pk.y = k0_.y; pk.z = k1_.x;
                                        ok = (int)(__double_as_longlong(k1_.y) >> 32) == sk;
                                    } else {
                                        int ak;
                                        pk = super_position(f, __ldg(f.idx3 + rowc + k), ak);
                                        ok = __ldg(f.spec + ak) == sk;
                                    }
                                }
                                double av[4], ad[4], bv[4], bd[4];
                                int ia = -1, ib = -1;
                                double dgk = 1.0;
                                if (ok) {
                                    ia = eval_leg_s(B, tk, dist_rn(pc, pk), av, ad);
                                    if (ia >= 0) {
                                        dgk = dist_rn(pg, pk);
                                        ib = eval_leg_s(B, tn, dgk, bv, bd);
                                    }
                                    ok = ia >= 0 && ib >= 0;
                                }
                                const unsigned live = __ballot_sync(FULL, ok);
                                if (ok) {
                                    const unsigned slot = (unsigned)__popc(live & ((1u << lane) - 1u));
                                    const unsigned r = rec_s + MR_REC * slot;
                                    sts128(r, make_double2(av[0], av[1]));
                                    sts128(r + 16, make_double2(av[2], av[3]));
                                    asm volatile("st.shared.s32 [%0], %1;" :: "r"(r + 32u), "r"(ia - tk.first) : "memory");
#pragma unroll
                                    for (int p = 0; p < 4; ++p) sts128(r + MR_A + 16u * p, make_double2(bv[p], bd[p]));
                                    asm volatile("st.shared.s32 [%0], %1;" :: "r"(r + MR_A + 64u), "r"(ib - tn.first) : "memory");
                                    if (role == 1) {
                                        const double inv = fast_rcp(dgk);
                                        sts128(r + MR_A + MR_B, make_double2((pk.x - pg.x) * inv, (pk.y - pg.y) * inv));
                                        sts64(r + MR_A + MR_B + 16, (pk.z - pg.z) * inv);
                                    }
                                }
                                __syncwarp();
                                // ---- contraction: every lane adds the live records to its cell
                                const int n_live = __popc(live);
#pragma unroll 2
                                for (int s = 0; s < n_live; ++s) {
                                    const unsigned r = rec_s + MR_REC * (unsigned)s;
                                    const unsigned qa = (unsigned)(cx - lds32(r + 32u)), qb = (unsigned)(cn - lds32(r + MR_A + 64u));
                                    const double x = lds64(qa < 4u ? r + 8u * qa : zero_s);
                                    if (role == 1) {
                                        const double2 w01 = lds128(r + MR_A + MR_B);
                                        const double w2 = lds64(r + MR_A + MR_B + 16);
                                        const double2 y = lds128(qb < 4u ? r + MR_A + 16u * qb : zero_s);
                                        const double t = x * y.y;
                                        P = fma(x, y.x, P);
                                        Q[0] = fma(w01.x, t, Q[0]);
                                        Q[1] = fma(w01.y, t, Q[1]);
                                        Q[2] = fma(w2, t, Q[2]);
                                    } else {
                                        P = fma(x, lds64(qb < 4u ? r + MR_A + 16u * qb : zero_s), P);
                                    }
                                }
                                __syncwarp();
                            }
                            any = true;
                            if (role == 0) add_own_leg<LA, 0>(R, gi - tg_.first, gv, gd, ug, P, Q, half_e);
                            else add_own_leg<LA, 1>(R, gi - tg_.first, gv, gd, ug, P, Q, half_e);
                        }

                        // ---- the kind's registers -> compressed columns of the lane's cell
                        if (!any) continue;
                        const int goff = __ldg(B.trio_goff + trio), col0 = __ldg(B.trio_col + trio);
                        const int dim_m = __ldg(B.trio_nk + 3 * trio + 1) - 4, dim_n = tn.nk - 4;
#pragma unroll
                        for (int yy = 0; yy < LA; ++yy) {
                            if (yy < tg_.count && has_cell) {
                                const int y = tg_.first + yy, x = tk.first + cx, n = tn.first + cn;
                                const int bin = goff + ((g_is_l ? y : x) * dim_m + (g_is_l ? x : y)) * dim_n + n;
                                const int col = __ldg(B.bin_col + bin);
                                if (col >= 0) {
                                    const double w = B.unit_weights ? 1.0 : __ldg(B.bin_w + bin);
                                    double2 *dst = reinterpret_cast<double2 *>(acc3 + 4 * (size_t)(col0 + col));
                                    double2 lo = dst[0], hi = dst[1];
                                    lo.x = fma(w, R[yy][0], lo.x);
                                    lo.y = fma(w, R[yy][1], lo.y);
                                    hi.x = fma(w, R[yy][2], hi.x);
                                    hi.y = fma(w, R[yy][3], hi.y);
                                    dst[0] = lo;
                                    dst[1] = hi;
                                }
                            }
                            __syncwarp();       // two y values of a lane can fold onto one column (symmetry 2)
                        }
                    }
                }
            }
        }

        // ---- rows fx_a, fy_a, fz_a
        __syncwarp();
        if (want_f) {
            for (int col = lane; col < F; col += 32) {
                double *src = (col < col3 ? acc : acc3) + 4 * (size_t)col;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    __stcs(xf + ((long long)c * f.n + a) * ld + col, src[1 + c]);      // written once: streaming
                    src[1 + c] = 0.0;
                }
            }
        }
        __syncwarp();
    }
    if (want_e)
        for (int col = lane; col < F; col += 32) partials[(size_t)gw * F + col] = (col < col3 ? acc : acc3)[4 * (size_t)col];
}

// Takes the frame when the basis needs this kernel: several species, a trio of symmetry 1, or 3-body
// rows longer than the unary kernels hold.  Returns 1 when it does not apply; error codes are <= 0.
int featurize_multi(uf3b_basis *basis, const uf3b_nlist *nl, double *x_energy, double *x_forces, int64_t ld,
                    cudaStream_t stream) {
    const BasisTab &T = basis->tab;
    if (T.n_trios < 1 || getenv("UF3B_NO_MULTI")) return 1;
    bool sym1 = false;
    int max_cells = 1, max_lm = 1;
    for (int t = 0; t < T.n_trios; ++t) {
        if (basis->h_trio_sym[t] == 3) return 1;        // three interchangeable legs: cells of one lane would collide
        if (basis->h_trio_sym[t] == 1) sym1 = true;
        const int trim = T.lead3 + T.trail3;
        const int l = basis->h_trio_dims[3 * t] - trim, m = basis->h_trio_dims[3 * t + 1] - trim;
        const int n = basis->h_trio_dims[3 * t + 2] - trim;
        max_cells = std::max(max_cells, std::max(l, m) * std::max(n, 1));
        max_lm = std::max(max_lm, std::max(l, m));
    }
    const bool wanted = T.ne > 1 || sym1 || nl->max3 > 32 || getenv("UF3B_MULTI");
    if (!wanted || max_cells > 128) return 1;
    const int F = basis->n_feats, n = (int)nl->n;
    int dev = 0, smem_max = 0, smem_sm = 0;
    UF3B_CUDA(cudaGetDevice(&dev));
    UF3B_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    UF3B_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    // small planes and rows the per-warp entry table holds: the kind-outer kernel with register accumulation
    const bool v2 = max_cells <= 32 && max_lm <= 8 && nl->max3 <= ME_MAX && !getenv("UF3B_MULTI_V1");
    MultiGeom mg = {};
    Multi2Geom mg2 = {};
    int col3 = F;
    for (int t = 0; t < T.n_trios; ++t) col3 = std::min(col3, basis->h_trio_col[t]);
    mg2.col3 = col3;
    mg.off_zero = (int)(((size_t)32 * F + 15) & ~size_t(15));
    mg2.off_zero = (int)(((size_t)32 * col3 + 15) & ~size_t(15));
    mg.off_rec = mg.off_zero + 16;
    mg2.off_ent = mg2.off_zero + 16;
    mg2.off_rec = mg2.off_ent + ME_MAX * (int)ME_REC;
    const size_t rec_bytes = std::max<size_t>(32 * MR_REC, 32 * sizeof(PairRec));
    mg.warp_bytes = (int)((mg.off_rec + rec_bytes + 15) & ~size_t(15));
    mg2.warp_bytes = (int)((mg2.off_rec + rec_bytes + 15) & ~size_t(15));
    mg2.n_knots3 = basis->n_knots3;
    mg2.n_poly3 = basis->n_poly3;
    mg2.off_warps = (int)((sizeof(double) * (size_t)(((basis->n_knots3 + 1) & ~1) + basis->n_poly3 / 16 * M2_PS) + 15) & ~size_t(15));
    const size_t tab_bytes = v2 ? (size_t)mg2.off_warps : 0;
    if (v2) mg.warp_bytes = mg2.warp_bytes;
    if (tab_bytes + (size_t)mg.warp_bytes > (size_t)smem_max) return 1;
    // warps per block x blocks per SM: the most resident warps that fit the shared memory of an SM
    int warps = 1, best = 0;
    for (int w = 4; w >= 1; --w) {
        const size_t blk = tab_bytes + (size_t)w * mg.warp_bytes;
        if (blk > (size_t)smem_max) continue;
        const int resident = std::min((int)((size_t)smem_sm / (blk + 1024)), 16 / w) * w;
        if (resident > best) { best = resident; warps = w; }
    }
    const size_t smem = tab_bytes + (size_t)warps * mg.warp_bytes;
    auto kernel = max_cells <= 32 ? k_rows_multi<1> : (max_cells <= 64 ? k_rows_multi<2> : k_rows_multi<4>);
    auto kernel2 = max_lm <= 4 ? k_rows_multi2<4> : k_rows_multi2<8>;
    const void *kfn = v2 ? (const void *)kernel2 : (const void *)kernel;
    UF3B_CUDA(ensure_dynamic_smem(kfn, smem));
    int per_sm = 1;
    if (v2) UF3B_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel2, warps * 32, smem));
    else UF3B_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, warps * 32, smem));
    if (per_sm < 1) per_sm = 1;
    int grid = std::max(1, sm_count() * per_sm / std::max(1, basis->frames_in_flight));
    grid = std::min(grid, (n + warps - 1) / warps);
    const int n_gw = grid * warps;

    const bool e_dev = x_energy && is_device_pointer(x_energy);
    const bool f_dev = x_forces && is_device_pointer(x_forces);
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
    if (v2) UF3B_CUDA(basis->gacc.reserve((size_t)n_gw * 4 * F));
    if (v2)
        UF3B_LAUNCH(kernel2, grid, warps * 32, smem, stream, basis->tab, view, mg2, d_xf, d_ld, basis->partials.p,
                    basis->gacc.p, x_energy ? 1 : 0, x_forces ? 1 : 0);
    else
        UF3B_LAUNCH(kernel, grid, warps * 32, smem, stream, basis->tab, view, mg, d_xf, d_ld, basis->partials.p,
                    x_energy ? 1 : 0, x_forces ? 1 : 0);
    if (g_timing) UF3B_CUDA(cudaEventRecord(ev1, stream));
    if (x_energy)
        if (int rc = launch_energy_row(basis->partials.p, n_gw, F, d_xe, stream)) return rc;
    return finish_featurize(basis, x_energy, x_forces, ld, d_xe, d_xf, F, n, e_dev, f_dev, stream, ev0, ev1);
}

}  // namespace uf3b
