// featurize.cu — Kernel B (fit path): fused pair + triplet B-spline feature rows.
//
// Replaces BasisFeaturizer.featurize_energy_2B/3B, featurize_force_2B/3B and the row
// assembly of evaluate_configuration (representation/process.py:293-506), i.e.
// bspline.evaluate_basis_functions / featurize_force_2B (bspline.py:810-895) and
// angles.featurize_energy_3b / featurize_force_3b with their numba scatters
// (angles.py:17-286) followed by compress_3B (bspline.py:664-690).
//
// One warp owns one real atom and produces that atom's three force rows (fx, fy, fz)
// plus its share of the energy row.  Accumulators live in shared memory
// ([column][e, fx, fy, fz]); nothing is scattered to other atoms' rows, so there are no
// atomics and the result is bit-reproducible run to run.
//   2-body: lanes evaluate 32 pairs at a time into shared records, then every lane
//           gathers the records that touch ITS feature column.
//   3-body: lanes evaluate 32 triangles at a time (three legs each) into shared records;
//           then the warp walks the records with lane = (p, q, r-pair) of the 4x4x4
//           block of non-zero basis products and adds them into the compressed column
//           of each bin.  Bins that fold onto the same column under the trio's
//           permutation symmetry are applied in separate phases.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "geom.cuh"
#include "spline.cuh"
#include "triangle.cuh"
#include "featurize_common.cuh"

namespace uf3b {

// One triangle as seen by the accumulating atom.  Four header quads first (broadcast
// 16-byte loads in phase B, no bit unpacking), then the leg values and role vectors.
struct __align__(16) TriRec {
    int base, mn, nn, col0;        // bin index of (il,im,in) incl. trio offset; bin strides
                                   // of l and m; first feature column of the trio
    int flags, sym, dlm, dmn;      // RF_*; symmetry order; il-im, im-in (mirror classes)
    int n_act, p0, q0, r0;         // untrimmed bins: count and first (p, q, r)
    int nq, nr, mq, mr;            // extents of q and r and their fixed-point reciprocals
    int il, im, in, pad0;          // first basis index of each leg (register-tile path)
    double v[3][4], dv[3][4];
    double A[3], B[3], C[3];
    double pad;
};
static_assert(sizeof(TriRec) % 16 == 0, "TriRec must keep 16-byte alignment in arrays");
constexpr int RF_VALID = 1, RF_CENTRE = 2;
constexpr unsigned REC_V = 80, REC_DV = 80 + 96, REC_ABC = 80 + 192;   // byte offsets

constexpr size_t WARP_SCRATCH = CHUNK * sizeof(TriRec) + sizeof(RoleViews);

// Per-warp shared memory: [accumulators (unless they live in global memory)][scratch]
__host__ __device__ inline size_t featurize_acc_bytes(int n_feats, bool global_acc) {
    return global_acc ? 0 : (((size_t)4 * n_feats * sizeof(double) + 15) & ~size_t(15));
}
__host__ __device__ inline size_t featurize_warp_bytes(int n_feats, bool global_acc) {
    return featurize_acc_bytes(n_feats, global_acc) + ((WARP_SCRATCH + 15) & ~size_t(15));
}

__device__ __forceinline__ void store_record(TriRec *rec, const Triangle &T, const BasisTab &B, int role) {
#pragma unroll
    for (int leg = 0; leg < 3; ++leg)
#pragma unroll
        for (int q = 0; q < 4; ++q) { rec->v[leg][q] = T.v[leg][q]; rec->dv[leg][q] = T.dv[leg][q]; }
#pragma unroll
    for (int c = 0; c < 3; ++c) { rec->A[c] = T.A[c]; rec->B[c] = T.B[c]; rec->C[c] = T.C[c]; }
    rec->mn = T.dim_m * T.dim_n;
    rec->nn = T.dim_n;
    rec->base = __ldg(B.trio_goff + T.trio) + (T.il * T.dim_m + T.im) * T.dim_n + T.in;
    rec->col0 = __ldg(B.trio_col + T.trio);
    rec->dlm = T.il - T.im;
    rec->dmn = T.im - T.in;
    rec->il = T.il;
    rec->im = T.im;
    rec->in = T.in;
    const int sym = __ldg(B.trio_sym + T.trio);
    rec->sym = sym < 1 ? 1 : (sym > 3 ? 3 : sym);
    // floor(b / n) for b < 64, n in 1..4 is (b * magic[n]) >> 8
    const int nq = T.cnt[1], nr = T.cnt[2];
    rec->nq = nq;
    rec->nr = nr;
    rec->mq = nq == 3 ? 86 : (256 >> (nq >> 1));
    rec->mr = nr == 3 ? 86 : (256 >> (nr >> 1));
    rec->n_act = T.cnt[0] * nq * nr;
    rec->p0 = T.lo[0];
    rec->q0 = T.lo[1];
    rec->r0 = T.lo[2];
    rec->flags = RF_VALID | (role == 0 ? RF_CENTRE : 0);
}

// Where a warp's accumulators [column][e, fx, fy, fz] live.
struct SharedAcc {
    unsigned base;       // shared-window byte address
    __device__ __forceinline__ void add(int col, double e, double x, double y, double z) const {
        const unsigned a = base + 32u * (unsigned)col;
        double2 u = lds128(a), w = lds128(a + 16);
        u.x += e; u.y += x; w.x += y; w.y += z;
        sts128(a, u);
        sts128(a + 16, w);
    }
};
struct GlobalAcc {
    double *base;
    __device__ __forceinline__ void add(int col, double e, double x, double y, double z) const {
        double2 *d = reinterpret_cast<double2 *>(base + 4 * (size_t)col);
        double2 u = d[0], w = d[1];
        u.x += e; u.y += x; w.x += y; w.y += z;
        d[0] = u;
        d[1] = w;
    }
};

// Phase B: scatter `count` records into the warp's accumulators.  A record whose untrimmed
// part of the 4x4x4 block of basis products holds at most 32 bins gets one bin per lane
// (bins enumerated by fixed-point division); otherwise lane = (p, q, r-pair) with two bins
// per lane.  Bins that fold onto the same compressed column under the trio's permutation
// symmetry are applied in separate passes, and only lanes with a kept bin of the pass
// touch the accumulators.
template <class Acc>
__device__ __forceinline__ void scatter_records(const BasisTab &B, unsigned recs, int count,
                                                const Acc acc, int lane, bool want_e) {
    for (int t = 0; t < count; ++t) {
        const unsigned rec = recs + (unsigned)t * (unsigned)sizeof(TriRec);
        const int4 h1 = lds128i(rec + 16);      // flags, sym, dlm, dmn
        if (!(h1.x & RF_VALID)) continue;
        const int4 h0 = lds128i(rec);           // base, mn, nn, col0
        const int4 h2 = lds128i(rec + 32);      // n_act, p0, q0, r0
        const bool centre = (h1.x & RF_CENTRE) != 0 && want_e;
        const int sym = h1.y;
        int p, q, r;
        const bool wide = h2.x > 32;
        if (!wide) {
            const int4 h3 = lds128i(rec + 48);  // nq, nr, mq, mr
            const int t1 = (lane * h3.w) >> 8;
            r = h2.w + lane - t1 * h3.y;
            const int t2 = (t1 * h3.z) >> 8;
            q = h2.z + t1 - t2 * h3.x;
            p = h2.y + t2;
        } else {
            p = lane >> 3;
            q = (lane >> 1) & 3;
            r = (lane & 1) * 2;
        }
        const int bin = h0.x + p * h0.y + q * h0.z + r;
        const bool in0 = wide || lane < h2.x;
        const int col_a = in0 ? __ldg(B.bin_col + bin) : -1;
        const int col_b = wide ? __ldg(B.bin_col + bin + 1) : -1;
        const bool live0 = col_a >= 0, live1 = col_b >= 0;
        double e0 = 0.0, e1 = 0.0, a0[3], a1[3];
        if (live0 || live1) {
            const double vl = lds64(rec + REC_V + 8 * p), dvl = lds64(rec + REC_DV + 8 * p);
            const double vm = lds64(rec + REC_V + 32 + 8 * q), dvm = lds64(rec + REC_DV + 32 + 8 * q);
            const double ga = dvl * vm, gb = vl * dvm, gc = vl * vm;
            const double2 ab0 = lds128(rec + REC_ABC), ab1 = lds128(rec + REC_ABC + 16);
            const double2 ab2 = lds128(rec + REC_ABC + 32), ab3 = lds128(rec + REC_ABC + 48);
            const double c2 = lds64(rec + REC_ABC + 64);
            // A = (ab0.x, ab0.y, ab1.x)  B = (ab1.y, ab2.x, ab2.y)  C = (ab3.x, ab3.y, c2)
            const double q1x = ga * ab0.x + gb * ab1.y, q2x = gc * ab3.x;
            const double q1y = ga * ab0.y + gb * ab2.x, q2y = gc * ab3.y;
            const double q1z = ga * ab1.x + gb * ab2.y, q2z = gc * c2;
            if (!wide) {
                const double vn = lds64(rec + REC_V + 64 + 8 * r), dvn = lds64(rec + REC_DV + 64 + 8 * r);
                e0 = centre ? gc * vn : 0.0;
                a0[0] = vn * q1x + dvn * q2x;
                a0[1] = vn * q1y + dvn * q2y;
                a0[2] = vn * q1z + dvn * q2z;
                a1[0] = a1[1] = a1[2] = 0.0;
            } else {
                const double2 vn = lds128(rec + REC_V + 64 + 8 * r), dvn = lds128(rec + REC_DV + 64 + 8 * r);
                e0 = centre ? gc * vn.x : 0.0;
                e1 = centre ? gc * vn.y : 0.0;
                a0[0] = vn.x * q1x + dvn.x * q2x;
                a0[1] = vn.x * q1y + dvn.x * q2y;
                a0[2] = vn.x * q1z + dvn.x * q2z;
                a1[0] = vn.y * q1x + dvn.y * q2x;
                a1[1] = vn.y * q1y + dvn.y * q2y;
                a1[2] = vn.y * q1z + dvn.y * q2z;
            }
            if (!B.unit_weights) {
                const double w0 = live0 ? __ldg(B.bin_w + bin) : 0.0;
                const double w1 = live1 ? __ldg(B.bin_w + bin + 1) : 0.0;
                e0 *= w0; e1 *= w1;
#pragma unroll
                for (int c = 0; c < 3; ++c) { a0[c] *= w0; a1[c] *= w1; }
            }
        }
        auto pass = [&](bool mine0, bool mine1) {
            if (mine0) acc.add(h0.w + col_a, e0, a0[0], a0[1], a0[2]);
            if (mine1) acc.add(h0.w + col_b, e1, a1[0], a1[1], a1[2]);
            __syncwarp();
        };
        if (sym == 1) {
            pass(live0, live1);
        } else if (sym == 2) {
            // (l,m,n) and (m,l,n) share a column: l <= m first, then l > m
            const bool upper = (p + h1.z) > q;
            pass(live0 && !upper, live1 && !upper);
            if (__any_sync(FULL, (live0 || live1) && upper)) pass(live0 && upper, live1 && upper);
        } else {
            // stable-sort class of (il+p, im+q, in+r): bins of one class never share a column
            const int x = p + h1.z + h1.w, y = q + h1.w;      // l, m relative to `in`
            const int cls0 = (x > y) | ((y > r) << 1) | ((x > r) << 2);
            const int cls1 = (x > y) | ((y > r + 1) << 1) | ((x > r + 1) << 2);
            for (int ph = 0; ph < 8; ++ph) {
                const bool m0 = live0 && cls0 == ph, m1 = live1 && cls1 == ph;
                if (__any_sync(FULL, m0 || m1)) pass(m0, m1);
            }
        }
    }
}

// ---------------------------------------------------------------- register-tile path
// For a unary basis whose UNTRIMMED 3-body grid is small (the default and demo UF3 bases:
// 3 x 3 x 9 cells after the trims), the per-atom grid is kept in registers instead of being
// updated in shared memory per triangle: lane owns up to KP (m, n) cells of the untrimmed
// (m, n) plane and, for each, one accumulator quad per untrimmed l (at most RT_LA).
// Phase A writes every leg's four values DENSELY by absolute basis index (zero outside the
// triangle's block), so phase B is branch-free: fixed per-lane offsets, a few broadcast
// loads and FMAs per record, no read-modify-write, no symmetry passes, no synchronisation.
// The grid is folded into the compressed columns (bin_col, mirror passes) once per atom.
constexpr int RT_LA = 4;
constexpr unsigned TR_ABC = 16, TR_VL = 96, TR_DVL = 128, TR_LEGS = 160;   // tile record bytes
static_assert(TR_LEGS + 16 * 12 <= sizeof(TriRec), "tile record must fit the scratch slot");

struct TileGeom {
    int l0, m0, n0;        // first untrimmed basis index per leg
    int la, ma, na;        // untrimmed extents (la <= RT_LA, ma + na <= 12, ma * na <= 64)
    int dim_m, dim_n;      // full grid extents of legs m, n
    int goff, col0, sym;
    int warp_bytes;        // shared memory per warp (accumulators + scratch)
    const unsigned char *leg_cache;   // k_leg_cache records (cached leg-grouped path) or null
    int cache_max3, cache_stride;     // L-record slots and records per centre
};

// Tile record: [flags][A B C][vl[4] dvl[4]][vm[ma] dvm[ma] vn[na] dvn[na]], dense by
// (basis index - first untrimmed index).
__device__ __forceinline__ void store_tile_record(unsigned char *rec, const Triangle &T,
                                                  const TileGeom &g, int role) {
    double *legs = reinterpret_cast<double *>(rec + TR_VL);
    {   // clear the dense leg arrays, 16 bytes at a time
        double2 *z = reinterpret_cast<double2 *>(legs);
        const int n16 = 4 + g.ma + g.na;
        for (int k = 0; k < n16; ++k) z[k] = make_double2(0.0, 0.0);
    }
    double *abc = reinterpret_cast<double *>(rec + TR_ABC);
#pragma unroll
    for (int c = 0; c < 3; ++c) { abc[c] = T.A[c]; abc[3 + c] = T.B[c]; abc[6 + c] = T.C[c]; }
    double *vm = legs + 8, *dvm = vm + g.ma, *vn = dvm + g.ma, *dvn = vn + g.na;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int l = T.il + p - g.l0, m = T.im + p - g.m0, n = T.in + p - g.n0;
        if (l >= 0 && l < g.la) { legs[l] = T.v[0][p]; legs[4 + l] = T.dv[0][p]; }
        if (m >= 0 && m < g.ma) { vm[m] = T.v[1][p]; dvm[m] = T.dv[1][p]; }
        if (n >= 0 && n < g.na) { vn[n] = T.v[2][p]; dvn[n] = T.dv[2][p]; }
    }
    *reinterpret_cast<int *>(rec) = RF_VALID | (role == 0 ? RF_CENTRE : 0);
}

template <int KP>
struct Tile {
    double acc[KP][RT_LA][4];
    int m[KP], n[KP];          // owned (m, n) cells (absolute basis indices), m < 0: none
    unsigned off_m[KP], off_n[KP];   // byte offsets of vm[m - m0], vn[n - n0] in a tile record

    __device__ __forceinline__ void init(const TileGeom &g, int lane) {
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            const int cell = lane + 32 * k;
            const bool ok = cell < g.ma * g.na;
            const int mi = ok ? cell / g.na : 0, ni = ok ? cell % g.na : 0;
            m[k] = ok ? g.m0 + mi : -1000;
            n[k] = ok ? g.n0 + ni : -1000;
            off_m[k] = TR_LEGS + 8u * (unsigned)mi;
            off_n[k] = TR_LEGS + 8u * (unsigned)(2 * g.ma + ni);
#pragma unroll
            for (int l = 0; l < RT_LA; ++l)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[k][l][c] = 0.0;
        }
    }

    // add `count` records (phase B of the register-tile path)
    __device__ __forceinline__ void accumulate(const TileGeom &g, unsigned recs, int count, bool want_e) {
        const unsigned d_m = 8u * (unsigned)g.ma, d_n = 8u * (unsigned)g.na;
        for (int t = 0; t < count; ++t) {
            const unsigned rec = recs + (unsigned)t * (unsigned)sizeof(TriRec);
            const int4 h = lds128i(rec);
            if (!(h.x & RF_VALID)) continue;
            const double escale = ((h.x & RF_CENTRE) != 0 && want_e) ? 1.0 : 0.0;
            const double2 ab0 = lds128(rec + TR_ABC), ab1 = lds128(rec + TR_ABC + 16);
            const double2 ab2 = lds128(rec + TR_ABC + 32), ab3 = lds128(rec + TR_ABC + 48);
            const double c2 = lds64(rec + TR_ABC + 64);
            // A = (ab0.x, ab0.y, ab1.x)  B = (ab1.y, ab2.x, ab2.y)  C = (ab3.x, ab3.y, c2)
            const double2 vl01 = lds128(rec + TR_VL), vl23 = lds128(rec + TR_VL + 16);
            const double2 dl01 = lds128(rec + TR_DVL), dl23 = lds128(rec + TR_DVL + 16);
            const double vl[4] = {vl01.x, vl01.y, vl23.x, vl23.y};
            const double dvl[4] = {dl01.x, dl01.y, dl23.x, dl23.y};
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                const double vm = lds64(rec + off_m[k]), dvm = lds64(rec + off_m[k] + d_m);
                const double vn = lds64(rec + off_n[k]), dvn = lds64(rec + off_n[k] + d_n);
                const double t1 = vm * vn, t2 = dvm * vn, t3 = vm * dvn, t1e = t1 * escale;
#pragma unroll
                for (int l = 0; l < RT_LA; ++l) {
                    if (l < g.la) {           // warp-uniform
                        const double ga = dvl[l] * t1, gb = vl[l] * t2, gc = vl[l] * t3;
                        acc[k][l][0] = fma(vl[l], t1e, acc[k][l][0]);
                        acc[k][l][1] = fma(gc, ab3.x, fma(gb, ab1.y, fma(ga, ab0.x, acc[k][l][1])));
                        acc[k][l][2] = fma(gc, ab3.y, fma(gb, ab2.x, fma(ga, ab0.y, acc[k][l][2])));
                        acc[k][l][3] = fma(gc, c2, fma(gb, ab2.y, fma(ga, ab1.x, acc[k][l][3])));
                    }
                }
            }
        }
    }

    // fold the grid into the compressed columns (once per atom) and clear it
    template <class Acc>
    __device__ __forceinline__ void flush(const BasisTab &B, const TileGeom &g, const Acc out) {
#pragma unroll
        for (int k = 0; k < KP; ++k)
#pragma unroll
            for (int l = 0; l < RT_LA; ++l) {
                const int ll = g.l0 + l;
                int col = -1;
                if (l < g.la && m[k] >= 0) col = __ldg(B.bin_col + g.goff + (ll * g.dim_m + m[k]) * g.dim_n + n[k]);
                const bool live = col >= 0;
                // stable-sort class of (l, m, n): equal classes never share a column
                int cls = 0;
                if (g.sym >= 2) cls = ll > m[k];
                if (g.sym == 3) cls |= ((m[k] > n[k]) << 1) | ((ll > n[k]) << 2);
                const int n_cls = g.sym == 1 ? 1 : (g.sym == 2 ? 2 : 8);
                for (int ph = 0; ph < n_cls; ++ph) {
                    if (live && cls == ph)
                        out.add(g.col0 + col, acc[k][l][0], acc[k][l][1], acc[k][l][2], acc[k][l][3]);
                    __syncwarp();
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[k][l][c] = 0.0;
            }
    }
};

// ---------------------------------------------------------------- leg-grouped tile path
// For a unary basis whose trio is symmetric in its two centre legs (symmetry >= 2: the l
// and m legs share knots, so which neighbour is called l does not change the folded column),
// the 3-body rows factor per LEG instead of per triangle:
//   centre role    x_a += sum_j  u_aj  dB(r_aj)[l] * P_j[m,n],   P_j = sum_{k != j} B(r_ak)[m] B(r_jk)[n]
//   energy         e   += sum_j        B(r_aj)[l]  * Pe_j[m,n],  Pe_j = same sum over k above j
//   neighbour role x_a += -u_ia dB(r_ia)[l] * P[m,n] + B(r_ia)[l] * Q[m,n],
//                  P = sum_k B(r_ik)[m] B(r_ak)[n],   Q = sum_k w_ak B(r_ik)[m] dB(r_ak)[n]
// so the inner loop over partners k updates one or four scalars per lane (lane = (m, n) cell)
// and the l dimension only appears once per leg group.  Phase A evaluates LEGS (one per
// lane, dense by absolute basis index) instead of triangles: each leg is shared by all the
// triangles of its group.  Requires every 3-body row involved to hold at most 32 entries.
constexpr unsigned LG_LM = 96;          // leg (centre, x): V[4] DV[4] u[3] pad
constexpr unsigned LG_N = 240;          // leg (x, y):      VN[12] DVN[12] w[3] pad
constexpr unsigned LG_N_BASE = 32 * LG_LM;
static_assert(LG_N_BASE + 32 * LG_N <= CHUNK * sizeof(TriRec), "leg caches must fit the scratch area");

// One leg, dense: values / derivatives of the untrimmed basis functions [x0, x0+xa) at
// d = |to - from| (zero when the leg is outside its knot range: the reference drops the
// whole triangle, angles.py:502-508), and the unit vector from -> to.
__device__ __forceinline__ void eval_dense_leg(const BasisTab &B, int leg, const Vec3 &from, const Vec3 &to,
                                               int x0, int xa, int stride, unsigned char *out) {
    double *val = reinterpret_cast<double *>(out), *der = val + stride, *uv = der + stride;
    for (int k = 0; k < stride; ++k) { val[k] = 0.0; der[k] = 0.0; }
    const double d = dist_rn(from, to);
    const int nk = __ldg(B.trio_nk + leg);
    const double *t = B.knots3 + __ldg(B.trio_koff + leg);
    double inv = 0.0;
    if (d >= t[0] && d <= t[nk - 1]) {
        double v[4], dv[4];
        const int idx = eval_leg(t, nk, __ldg(B.trio_scale + leg), B.poly3 + __ldg(B.trio_poff + leg), d,
                                 B.lead3, B.trail3, v, dv);
        if (idx >= 0) {
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int x = idx + p - x0;
                if (x >= 0 && x < xa) { val[x] = v[p]; der[x] = dv[p]; }
            }
        }
        inv = fast_rcp(d);
    }
    uv[0] = (to.x - from.x) * inv;
    uv[1] = (to.y - from.y) * inv;
    uv[2] = (to.z - from.z) * inv;
}

__device__ __forceinline__ void legs_three_body(const BasisTab &B, const FrameView &f, const TileGeom &g,
                                                int a, const Vec3 &pa, unsigned char *scratch, unsigned scratch_s,
                                                Tile<1> &tile, int lane, bool want_e, bool want_f) {
    const int row0 = __ldg(f.off3 + a), n3a = __ldg(f.cnt3 + a);
    if (n3a < 1) return;
    const unsigned lm_s = scratch_s, nn_s = scratch_s + LG_N_BASE;
    unsigned char *lm = scratch, *nn = scratch + LG_N_BASE;
    const int cell = lane < g.ma * g.na ? lane : 0;
    const unsigned off_m = 8u * (unsigned)(cell / g.na), off_n = 8u * (unsigned)(cell % g.na);
    int dummy;

    // ---- (i) `a` as the centre
    if (lane < n3a)
        eval_dense_leg(B, 0, pa, super_position(f, __ldg(f.idx3 + row0 + lane), dummy), g.l0, g.la, 4,
                       lm + lane * LG_LM);
    __syncwarp();
    const int np = n3a - 1;                       // partners per group
    const int per_pass = np > 0 ? 32 / np : 32;   // groups per pass (n3a <= 32)
    for (int g0 = 0; g0 < n3a && np > 0; g0 += per_pass) {
        {   // phase A: legs (j, k) of up to per_pass groups, one per lane
            const int gi = lane / np, s = lane - gi * np, j = g0 + gi;
            if (gi < per_pass && j < n3a) {
                const int k = s + (s >= j);
                eval_dense_leg(B, 2, super_position(f, __ldg(f.idx3 + row0 + j), dummy),
                               super_position(f, __ldg(f.idx3 + row0 + k), dummy), g.n0, g.na, 12,
                               nn + lane * LG_N);
            }
        }
        __syncwarp();
        for (int gi = 0; gi < per_pass && g0 + gi < n3a; ++gi) {      // phase B: one group at a time
            const int j = g0 + gi;
            double P = 0.0, Pe = 0.0;
            for (int s = 0; s < np; ++s) {
                const int k = s + (s >= j);
                const double vm = lds64(lm_s + (unsigned)k * LG_LM + off_m);
                const double vn = lds64(nn_s + (unsigned)(gi * np + s) * LG_N + off_n);
                P = fma(vm, vn, P);
                if (k > j) Pe = fma(vm, vn, Pe);       // each unordered pair once for the energy row
            }
            const unsigned lj = lm_s + (unsigned)j * LG_LM;
            const double2 u01 = lds128(lj + 64);
            const double u2 = lds64(lj + 80);
#pragma unroll
            for (int l = 0; l < RT_LA; ++l) {
                if (l < g.la) {
                    const double v = lds64(lj + 8 * l), dP = lds64(lj + 32 + 8 * l) * P;
                    if (want_e) tile.acc[0][l][0] = fma(v, Pe, tile.acc[0][l][0]);
                    tile.acc[0][l][1] = fma(u01.x, dP, tile.acc[0][l][1]);
                    tile.acc[0][l][2] = fma(u01.y, dP, tile.acc[0][l][2]);
                    tile.acc[0][l][3] = fma(u2, dP, tile.acc[0][l][3]);
                }
            }
        }
        __syncwarp();
    }
    if (!want_f) return;

    // ---- (ii) `a` as a neighbour of every centre i in its list
    for (int e = 0; e < n3a; ++e) {
        const int m = __ldg(f.idx3 + row0 + e);
        const int gimg = image_of(f, m);
        const int ci = m - gimg * f.n;
        const int apr = __ldg(f.img_inv + gimg) * f.n + a;
        const int rowi = __ldg(f.off3 + ci), ni = __ldg(f.cnt3 + ci);
        const int mine = lane < ni ? __ldg(f.idx3 + rowi + lane) : -1;
        const unsigned hit = __ballot_sync(FULL, mine == apr);
        if (!hit) continue;                       // one-ulp asymmetry of the list criterion
        const int qa = __ffs(hit) - 1;
        const Vec3 pi = real_position(f, ci), pap = super_position(f, apr, dummy);
        // phase A: legs (i, x) for the whole row of i, then legs (a', k) for k != a'
        const int n_items = 2 * ni - 1;
        for (int it0 = 0; it0 < n_items; it0 += 32) {
            const int it = it0 + lane;
            if (it < n_items) {                   // one call site: both leg kinds share the code
                const bool centre_leg = it < ni;
                const int s = it - ni, k = centre_leg ? it : s + (s >= qa);
                eval_dense_leg(B, centre_leg ? 0 : 2, centre_leg ? pi : pap,
                               super_position(f, __ldg(f.idx3 + rowi + k), dummy),
                               centre_leg ? g.l0 : g.n0, centre_leg ? g.la : g.na, centre_leg ? 4 : 12,
                               centre_leg ? lm + it * LG_LM : nn + s * LG_N);
            }
        }
        __syncwarp();
        double P = 0.0, Qx = 0.0, Qy = 0.0, Qz = 0.0;
        for (int s = 0; s < ni - 1; ++s) {
            const int k = s + (s >= qa);
            const unsigned ns = nn_s + (unsigned)s * LG_N;
            const double vm = lds64(lm_s + (unsigned)k * LG_LM + off_m);
            const double vn = lds64(ns + off_n), dvn = lds64(ns + 96 + off_n);
            const double2 w01 = lds128(ns + 192);
            const double w2 = lds64(ns + 208);
            P = fma(vm, vn, P);
            const double t3 = vm * dvn;
            Qx = fma(w01.x, t3, Qx);
            Qy = fma(w01.y, t3, Qy);
            Qz = fma(w2, t3, Qz);
        }
        const unsigned la_s = lm_s + (unsigned)qa * LG_LM;
        const double2 u01 = lds128(la_s + 64);
        const double u2 = lds64(la_s + 80);
#pragma unroll
        for (int l = 0; l < RT_LA; ++l) {
            if (l < g.la) {
                const double v = lds64(la_s + 8 * l), dP = lds64(la_s + 32 + 8 * l) * P;
                tile.acc[0][l][1] += v * Qx - u01.x * dP;
                tile.acc[0][l][2] += v * Qy - u01.y * dP;
                tile.acc[0][l][3] += v * Qz - u2 * dP;
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------- plane path
// The leg-grouped factorisation above for grids of ANY size (the manuscript basis: 7 x 7 x 17
// untrimmed cells, 456 columns): nothing is kept dense in registers.  Legs are stored
// sparsely (first basis index + 4 values + 4 derivatives + unit vector); for one leg group
// the partner sums P[m,n] (and, in the neighbour role, the vector sums Q[m,n]) are built in
// a per-warp shared-memory plane by lanes = (p, q) of the 4 x 4 non-zero products of each
// partner; then the plane is contracted with the group's 4 non-zero l values straight into
// the warp's column accumulators (bin_col gives the compressed column).
// One pass of the outer stage touches bins of ONE l, so two lanes never meet in a column
// for a trio of symmetry 2 ((l,m,n) ~ (m,l,n)); passes are ordered with __syncwarp.
// Lanes with nothing to add (trimmed / out-of-range basis index, empty cell, folded-away
// bin) aim their read-modify-write at a dummy slot instead of branching, so the loops are
// branch-free and the warp stays converged.
// Energy row: every unordered neighbour pair is met twice (as (j,k) and (k,j)) and both
// orders fold onto the same column, so e += (1/2) B_l(r_aj) P_j — exact halving.
constexpr unsigned SPL_REC = 96;        // l / m leg: v[4] dv[4] u[3] {int idx - x0, pad}
constexpr unsigned SPN_REC = 112;       // n leg:     v[4] dv[4] {1, wx, wy, wz} {int idx - x0, pad} pad
constexpr unsigned SPL_IDX = 88, SPN_IDX = 96;
constexpr unsigned SP_N_BASE = 32 * SPL_REC, SP_PLANE = SP_N_BASE + 32 * SPN_REC;
constexpr int PL_MAX_CELLS = 128;
constexpr int SP_DEAD = -(1 << 20);     // idx of a leg outside its knot range
__host__ __device__ inline size_t plane_scratch_bytes(int n_cells) {   // + one dummy quad
    return SP_PLANE + 32 * (size_t)n_cells + 32;
}

template <bool N_LEG>
__device__ __forceinline__ void eval_sparse_leg(const BasisTab &B, int leg, const Vec3 &from, const Vec3 &to,
                                                int x0, unsigned char *out) {
    double v[4] = {0.0, 0.0, 0.0, 0.0}, dv[4] = {0.0, 0.0, 0.0, 0.0};
    const double d = dist_rn(from, to);
    const int nk = __ldg(B.trio_nk + leg);
    const double *t = B.knots3 + __ldg(B.trio_koff + leg);
    double inv = 0.0;
    int rel = SP_DEAD;
    if (d >= t[0] && d <= t[nk - 1]) {              // angles.py:502-508 drops the whole triangle
        const int idx = eval_leg(t, nk, __ldg(B.trio_scale + leg), B.poly3 + __ldg(B.trio_poff + leg), d,
                                 B.lead3, B.trail3, v, dv);
        if (idx >= 0) rel = idx - x0;
        inv = fast_rcp(d);
    }
    const double ux = (to.x - from.x) * inv, uy = (to.y - from.y) * inv, uz = (to.z - from.z) * inv;
    const double tag = __longlong_as_double((long long)(unsigned)rel);
    double2 *o = reinterpret_cast<double2 *>(out);
    o[0] = make_double2(v[0], v[1]);
    o[1] = make_double2(v[2], v[3]);
    o[2] = make_double2(dv[0], dv[1]);
    o[3] = make_double2(dv[2], dv[3]);
    if (N_LEG) {
        o[4] = make_double2(1.0, ux);
        o[5] = make_double2(uy, uz);
        o[6] = make_double2(tag, 0.0);
    } else {
        o[4] = make_double2(ux, uy);
        o[5] = make_double2(uz, tag);
    }
}

// Contract the plane of one leg group with the group's own leg (record `lrec`) into the
// column accumulators and clear the plane.  CENTRE: slots 0/1 hold the partner sum P split
// over even / odd partners; else the slots are (P, Qx, Qy, Qz).
template <bool CENTRE, int KC>
__device__ __forceinline__ void plane_outer(const BasisTab &B, const TileGeom &g, unsigned plane_s, unsigned dummy_s,
                                            unsigned lrec, const int (&bin0)[KC], int n_cells, unsigned acc_s,
                                            int lane, bool want_e) {
    const int il = lds32(lrec + SPL_IDX);
    const double2 v01 = lds128(lrec), v23 = lds128(lrec + 16), d01 = lds128(lrec + 32), d23 = lds128(lrec + 48);
    const double2 u01 = lds128(lrec + 64);
    const double u2 = lds64(lrec + 80);
    const double v[4] = {v01.x, v01.y, v23.x, v23.y}, dv[4] = {d01.x, d01.y, d23.x, d23.y};
    const int lmn = g.dim_m * g.dim_n;
    const double half_e = want_e ? 0.5 : 0.0;
    double P[KC], Qx[KC], Qy[KC], Qz[KC];
    bool live[KC];
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
        const int cell = lane + 32 * kk;
        live[kk] = cell < n_cells;
        const unsigned pa = live[kk] ? plane_s + 32u * (unsigned)cell : dummy_s;
        const double2 a = lds128(pa);
        sts128(pa, make_double2(0.0, 0.0));
        if (CENTRE) {
            P[kk] = a.x + a.y;
            Qx[kk] = Qy[kk] = Qz[kk] = 0.0;
            live[kk] = live[kk] && P[kk] != 0.0;
        } else {
            const double2 b = lds128(pa + 16);
            sts128(pa + 16, make_double2(0.0, 0.0));
            P[kk] = a.x; Qx[kk] = a.y; Qy[kk] = b.x; Qz[kk] = b.y;
            live[kk] = live[kk] && (a.x != 0.0 || a.y != 0.0 || b.x != 0.0 || b.y != 0.0);
        }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int lrel = il + p;
        if ((unsigned)lrel < (unsigned)g.la) {       // warp-uniform
            unsigned ad[KC];
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) {
                const int col = live[kk] ? __ldg(B.bin_col + bin0[kk] + lrel * lmn) : -1;
                ad[kk] = col >= 0 ? acc_s + 32u * (unsigned)(g.col0 + col) : dummy_s;
            }
            double2 q0[KC], q1[KC];
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) { q0[kk] = lds128(ad[kk]); q1[kk] = lds128(ad[kk] + 16); }
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) {
                const double dP = dv[p] * P[kk];
                if (CENTRE) {
                    q0[kk].x = fma(half_e * v[p], P[kk], q0[kk].x);
                    q0[kk].y = fma(u01.x, dP, q0[kk].y);
                    q1[kk].x = fma(u01.y, dP, q1[kk].x);
                    q1[kk].y = fma(u2, dP, q1[kk].y);
                } else {
                    q0[kk].y += v[p] * Qx[kk] - u01.x * dP;
                    q1[kk].x += v[p] * Qy[kk] - u01.y * dP;
                    q1[kk].y += v[p] * Qz[kk] - u2 * dP;
                }
                sts128(ad[kk], q0[kk]);
                sts128(ad[kk] + 16, q1[kk]);
            }
            __syncwarp();
        }
    }
}

template <int KC>
__device__ __forceinline__ void plane_three_body(const BasisTab &B, const FrameView &f, const TileGeom &g,
                                                 int a, const Vec3 &pa, unsigned char *scratch, unsigned scratch_s,
                                                 unsigned acc_s, const int (&bin0)[KC], int lane,
                                                 bool want_e, bool want_f) {
    const int row0 = __ldg(f.off3 + a), n3a = __ldg(f.cnt3 + a);
    if (n3a < 1) return;
    const int n_cells = g.ma * g.na;
    const unsigned lm_s = scratch_s, nn_s = scratch_s + SP_N_BASE, plane_s = scratch_s + SP_PLANE;
    const unsigned dummy_s = plane_s + 32u * (unsigned)n_cells;
    unsigned char *lm = scratch, *nn = scratch + SP_N_BASE;
    const unsigned ma = (unsigned)g.ma, na = (unsigned)g.na;
    const int half = lane >> 4, p = (lane >> 2) & 3, q = lane & 3;
    int dummy;

    // ---- (i) `a` as the centre
    if (lane < n3a)
        eval_sparse_leg<false>(B, 0, pa, super_position(f, __ldg(f.idx3 + row0 + lane), dummy), g.l0,
                               lm + lane * SPL_REC);
    __syncwarp();
    const int per_pass = 32 / n3a;                // groups per pass (n3a <= 32); slot = gi * n3a + k
    for (int g0 = 0; g0 < n3a && n3a > 1; g0 += per_pass) {
        {   // phase A: legs (j, k) of up to per_pass groups, one per lane
            const int gi = lane / n3a, k = lane - gi * n3a, j = g0 + gi;
            if (gi < per_pass && j < n3a && k != j)
                eval_sparse_leg<true>(B, 2, super_position(f, __ldg(f.idx3 + row0 + j), dummy),
                                      super_position(f, __ldg(f.idx3 + row0 + k), dummy), g.n0,
                                      nn + lane * SPN_REC);
        }
        __syncwarp();
        for (int gi = 0; gi < per_pass && g0 + gi < n3a; ++gi) {
            const int j = g0 + gi;
            const unsigned n_row = nn_s + (unsigned)(gi * n3a) * SPN_REC;
            for (int k0 = 0; k0 < n3a; k0 += 2) {         // two partners per pass, one per half-warp
                const int k = min(k0 + half, n3a - 1);
                const unsigned lk = lm_s + (unsigned)k * SPL_REC, ns = n_row + (unsigned)k * SPN_REC;
                const unsigned mrel = (unsigned)(lds32(lk + SPL_IDX) + p);
                const unsigned nrel = (unsigned)(k == j ? SP_DEAD : lds32(ns + SPN_IDX) + q);
                const bool ok = mrel < ma && nrel < na && k0 + half < n3a;
                const unsigned ad = ok ? plane_s + 32u * (mrel * na + nrel) + 8u * (unsigned)half : dummy_s;
                sts64(ad, fma(lds64(lk + 8 * p), lds64(ns + 8 * q), lds64(ad)));
                __syncwarp();
            }
            plane_outer<true, KC>(B, g, plane_s, dummy_s, lm_s + (unsigned)j * SPL_REC, bin0, n_cells, acc_s,
                                  lane, want_e);
        }
        __syncwarp();
    }
    if (!want_f) return;

    // ---- (ii) `a` as a neighbour of every centre i in its list
    const unsigned off_a = 32u * (unsigned)half + 8u * (unsigned)q;   // v[q] | dv[q]
    const unsigned off_w = 64u + 16u * (unsigned)half;                // (1, wx) | (wy, wz)
    for (int e = 0; e < n3a; ++e) {
        const int m = __ldg(f.idx3 + row0 + e);
        const int gimg = image_of(f, m);
        const int ci = m - gimg * f.n;
        const int apr = __ldg(f.img_inv + gimg) * f.n + a;
        const int rowi = __ldg(f.off3 + ci), ni = __ldg(f.cnt3 + ci);
        const int mine = lane < ni ? __ldg(f.idx3 + rowi + lane) : -1;
        const unsigned hit = __ballot_sync(FULL, mine == apr);
        if (!hit) continue;                       // one-ulp asymmetry of the list criterion
        const int qa = __ffs(hit) - 1;
        const Vec3 pi = real_position(f, ci), pap = super_position(f, apr, dummy);
        // phase A: legs (i, x) for the whole row of i, then legs (a', k) for k != a' (slot k)
        for (int it0 = 0; it0 < 2 * ni; it0 += 32) {
            const int it = it0 + lane;
            const bool centre_leg = it < ni;
            const int k = centre_leg ? it : it - ni;
            if (it < 2 * ni && (centre_leg || k != qa)) {
                const Vec3 pk = super_position(f, __ldg(f.idx3 + rowi + k), dummy);
                if (centre_leg) eval_sparse_leg<false>(B, 0, pi, pk, g.l0, lm + k * SPL_REC);
                else eval_sparse_leg<true>(B, 2, pap, pk, g.n0, nn + k * SPN_REC);
            }
        }
        __syncwarp();
        for (int k = 0; k < ni; ++k) {            // lanes 0-15 own (P, Qx), lanes 16-31 (Qy, Qz)
            if (k == qa) continue;
            const unsigned lk = lm_s + (unsigned)k * SPL_REC, ns = nn_s + (unsigned)k * SPN_REC;
            const unsigned mrel = (unsigned)(lds32(lk + SPL_IDX) + p), nrel = (unsigned)(lds32(ns + SPN_IDX) + q);
            const bool ok = mrel < ma && nrel < na;
            const unsigned ad = ok ? plane_s + 32u * (mrel * na + nrel) + 16u * (unsigned)half : dummy_s;
            const double vm = lds64(lk + 8 * p), x = lds64(ns + off_a), dvn = lds64(ns + 32 + 8 * q);
            const double2 w = lds128(ns + off_w);
            double2 c = lds128(ad);
            c.x = fma(vm * x, w.x, c.x);          // P += vm vn       | Qy += wy vm dvn
            c.y = fma(vm * dvn, w.y, c.y);        // Qx += wx vm dvn  | Qz += wz vm dvn
            sts128(ad, c);
            __syncwarp();
        }
        plane_outer<false, KC>(B, g, plane_s, dummy_s, lm_s + (unsigned)qa * SPL_REC, bin0, n_cells, acc_s,
                               lane, false);
        __syncwarp();
    }
}

// ---------------------------------------------------------------- leg cache
// Every leg of the 3-body terms is shared: the leg (i, x) of centre i by all triangles of i
// through x, the leg (j, k) between two neighbours of i by centre i and by j and k in their
// neighbour roles.  The owner-computes kernels above re-evaluate them per visiting atom
// (574 leg evaluations per atom of bulk W against 105 distinct ones, 40 % of the
// instructions of the demo-basis kernel).  k_leg_cache evaluates each leg ONCE, in the frame
// of its real centre — the same positions and arithmetic the visiting atoms used, so the
// values are bit-identical — into sparse 96-byte records in global memory (L2-resident:
// 10 KB per atom); the cached leg-grouped path then only loads and densifies records.
// Layout per centre i: records [0, max3) = legs (i, row entry), then the pair (j < k by row
// position) at max3 + k (k - 1) / 2 + j, stored from j to k (unit vector j -> k).
__global__ void __launch_bounds__(128)
k_leg_cache(const BasisTab B, const FrameView f, const TileGeom g, int max3, int stride, unsigned char *cache) {
    // block = one centre, thread = one leg (105 per W atom): the evaluation is a chain of
    // dependent gathers (list entry -> position -> knots -> piece), so the parallelism has to
    // come from many short threads rather than from a warp looping over an atom's legs
    int dummy;
    for (int i = blockIdx.x; i < f.n; i += gridDim.x) {
        const int row = __ldg(f.off3 + i), ni = __ldg(f.cnt3 + i);
        unsigned char *mine = cache + (size_t)i * stride * SPL_REC;
        const int n_legs = ni + ni * (ni - 1) / 2;
        for (int t = threadIdx.x; t < n_legs; t += blockDim.x) {
            if (t < ni) {
                eval_sparse_leg<false>(B, 0, real_position(f, i), super_position(f, __ldg(f.idx3 + row + t), dummy),
                                       g.l0, mine + (size_t)t * SPL_REC);
            } else {
                int qj, qk;
                unrank_pair(t - ni, qj, qk);
                eval_sparse_leg<false>(B, 2, super_position(f, __ldg(f.idx3 + row + qj), dummy),
                                       super_position(f, __ldg(f.idx3 + row + qk), dummy), g.n0,
                                       mine + (size_t)(max3 + t - ni) * SPL_REC);
            }
        }
    }
}

// Cached sparse record (global) -> dense leg arrays in shared memory: `stride` values,
// `stride` derivatives, unit vector (times `sign`), as eval_dense_leg writes them (stride
// even, `out` 16-byte aligned).  VALUES_ONLY skips the derivatives and the unit vector.
template <bool VALUES_ONLY>
__device__ __forceinline__ void load_dense_leg(const unsigned char *rec, int xa, int stride, double sign,
                                               unsigned char *out) {
    const double2 *gp = reinterpret_cast<const double2 *>(rec);
    const double2 v01 = __ldg(gp), v23 = __ldg(gp + 1);
    double2 d01 = make_double2(0.0, 0.0), d23 = d01, u01 = d01;
    const double2 u2t = __ldg(gp + 5);
    if (!VALUES_ONLY) { d01 = __ldg(gp + 2); d23 = __ldg(gp + 3); u01 = __ldg(gp + 4); }
    const int rel = (int)__double_as_longlong(u2t.y);
    double *val = reinterpret_cast<double *>(out), *der = val + stride, *uv = der + stride;
    double2 *z = reinterpret_cast<double2 *>(out);
    const int n16 = VALUES_ONLY ? stride / 2 : stride;
    for (int k = 0; k < n16; ++k) z[k] = make_double2(0.0, 0.0);
    const double v[4] = {v01.x, v01.y, v23.x, v23.y}, dv[4] = {d01.x, d01.y, d23.x, d23.y};
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int x = rel + p;
        if (x >= 0 && x < xa) {
            val[x] = v[p];
            if (!VALUES_ONLY) der[x] = dv[p];
        }
    }
    if (!VALUES_ONLY) {
        uv[0] = sign * u01.x;
        uv[1] = sign * u01.y;
        uv[2] = sign * u2t.x;
    }
}

// legs_three_body with every leg evaluation replaced by a load from the leg cache.
__device__ __forceinline__ void legs_three_body_cached(const BasisTab &B, const FrameView &f, const TileGeom &g,
                                                       const unsigned char *cache, int max3, int stride,
                                                       int a, unsigned char *scratch, unsigned scratch_s,
                                                       Tile<1> &tile, int lane, bool want_e, bool want_f) {
    const int row0 = __ldg(f.off3 + a), n3a = __ldg(f.cnt3 + a);
    if (n3a < 1) return;
    const unsigned lm_s = scratch_s, nn_s = scratch_s + LG_N_BASE;
    unsigned char *lm = scratch, *nn = scratch + LG_N_BASE;
    const int cell = lane < g.ma * g.na ? lane : 0;
    const unsigned off_m = 8u * (unsigned)(cell / g.na), off_n = 8u * (unsigned)(cell % g.na);
    const unsigned char *mine = cache + (size_t)a * stride * SPL_REC;

    // ---- (i) `a` as the centre
    if (lane < n3a) load_dense_leg<false>(mine + (size_t)lane * SPL_REC, g.la, 4, 1.0, lm + lane * LG_LM);
    __syncwarp();
    const int np = n3a - 1;                       // partners per group
    const int per_pass = np > 0 ? 32 / np : 32;   // groups per pass (n3a <= 32)
    for (int g0 = 0; g0 < n3a && np > 0; g0 += per_pass) {
        {   // phase A: legs (j, k) of up to per_pass groups, one per lane
            const int gi = lane / np, s = lane - gi * np, j = g0 + gi;
            if (gi < per_pass && j < n3a) {
                const int k = s + (s >= j);
                const int lo = j < k ? j : k, hi = j < k ? k : j;
                load_dense_leg<true>(mine + (size_t)(max3 + hi * (hi - 1) / 2 + lo) * SPL_REC, g.na, 12, 1.0,
                                     nn + lane * LG_N);       // only the values are used in this role
            }
        }
        __syncwarp();
        for (int gi = 0; gi < per_pass && g0 + gi < n3a; ++gi) {      // phase B: one group at a time
            const int j = g0 + gi;
            double P = 0.0, Pe = 0.0;
            for (int s = 0; s < np; ++s) {
                const int k = s + (s >= j);
                const double vm = lds64(lm_s + (unsigned)k * LG_LM + off_m);
                const double vn = lds64(nn_s + (unsigned)(gi * np + s) * LG_N + off_n);
                P = fma(vm, vn, P);
                if (k > j) Pe = fma(vm, vn, Pe);       // each unordered pair once for the energy row
            }
            const unsigned lj = lm_s + (unsigned)j * LG_LM;
            const double2 u01 = lds128(lj + 64);
            const double u2 = lds64(lj + 80);
#pragma unroll
            for (int l = 0; l < RT_LA; ++l) {
                if (l < g.la) {
                    const double v = lds64(lj + 8 * l), dP = lds64(lj + 32 + 8 * l) * P;
                    if (want_e) tile.acc[0][l][0] = fma(v, Pe, tile.acc[0][l][0]);
                    tile.acc[0][l][1] = fma(u01.x, dP, tile.acc[0][l][1]);
                    tile.acc[0][l][2] = fma(u01.y, dP, tile.acc[0][l][2]);
                    tile.acc[0][l][3] = fma(u2, dP, tile.acc[0][l][3]);
                }
            }
        }
        __syncwarp();
    }
    if (!want_f) return;

    // ---- (ii) `a` as a neighbour of every centre i in its list
    // lane e resolves entry e up front (centre, its row length, the position qa of `a` in its
    // row), so that the per-centre loop below starts with the record loads
    int my_ci = 0, my_ni = 0, my_qa = -1;
    if (lane < n3a) {
        const int m = __ldg(f.idx3 + row0 + lane);
        const int gimg = image_of(f, m);
        my_ci = m - gimg * f.n;
        const int apr = __ldg(f.img_inv + gimg) * f.n + a;
        const int rowi = __ldg(f.off3 + my_ci);
        my_ni = __ldg(f.cnt3 + my_ci);
        for (int k = 0; k < my_ni; ++k)
            if (__ldg(f.idx3 + rowi + k) == apr) my_qa = k;
    }
    for (int e = 0; e < n3a; ++e) {
        const int ci = __shfl_sync(FULL, my_ci, e), ni = __shfl_sync(FULL, my_ni, e);
        const int qa = __shfl_sync(FULL, my_qa, e);
        if (qa < 0) continue;                     // one-ulp asymmetry of the list criterion
        const unsigned char *theirs = cache + (size_t)ci * stride * SPL_REC;
        // phase A: legs (i, x) for the whole row of i, then legs (a', k) for k != a'
        const int n_items = 2 * ni - 1;
        for (int it0 = 0; it0 < n_items; it0 += 32) {
            const int it = it0 + lane;
            if (it < n_items) {
                if (it < ni) {
                    load_dense_leg<false>(theirs + (size_t)it * SPL_REC, g.la, 4, 1.0, lm + it * LG_LM);
                } else {
                    const int s = it - ni, k = s + (s >= qa);
                    const int lo = qa < k ? qa : k, hi = qa < k ? k : qa;
                    load_dense_leg<false>(theirs + (size_t)(max3 + hi * (hi - 1) / 2 + lo) * SPL_REC, g.na, 12,
                                          qa < k ? 1.0 : -1.0, nn + s * LG_N);      // unit vector a' -> k
                }
            }
        }
        __syncwarp();
        double P = 0.0, Qx = 0.0, Qy = 0.0, Qz = 0.0;
        for (int s = 0; s < ni - 1; ++s) {
            const int k = s + (s >= qa);
            const unsigned ns = nn_s + (unsigned)s * LG_N;
            const double vm = lds64(lm_s + (unsigned)k * LG_LM + off_m);
            const double vn = lds64(ns + off_n), dvn = lds64(ns + 96 + off_n);
            const double2 w01 = lds128(ns + 192);
            const double w2 = lds64(ns + 208);
            P = fma(vm, vn, P);
            const double t3 = vm * dvn;
            Qx = fma(w01.x, t3, Qx);
            Qy = fma(w01.y, t3, Qy);
            Qz = fma(w2, t3, Qz);
        }
        const unsigned la_s = lm_s + (unsigned)qa * LG_LM;
        const double2 u01 = lds128(la_s + 64);
        const double u2 = lds64(la_s + 80);
#pragma unroll
        for (int l = 0; l < RT_LA; ++l) {
            if (l < g.la) {
                const double v = lds64(la_s + 8 * l), dP = lds64(la_s + 32 + 8 * l) * P;
                tile.acc[0][l][1] += v * Qx - u01.x * dP;
                tile.acc[0][l][2] += v * Qy - u01.y * dP;
                tile.acc[0][l][3] += v * Qz - u2 * dP;
            }
        }
        __syncwarp();
    }
}

// GLOBAL_ACC: the per-warp accumulators [4 * n_feats] live in a global scratch buffer
// (L1/L2 resident) instead of shared memory — the path for bases whose rows do not fit
// (e.g. 18 trio interactions of a ternary system, F ~ 7000).
// KP = 1, 2 selects the register-tile path with KP (m, n) cells per lane; KP = 3 the
// leg-grouped tile path (KP = 7: with the leg cache); KP = 4, 5, 6 the plane path with
// 1, 2, 4 chunks of 32 (m, n) cells.
template <bool GLOBAL_ACC, int KP>
__global__ void __launch_bounds__(256, 2)
k_featurize(const BasisTab B, const FrameView f, const TileGeom tg, double *__restrict__ xf, long long ld,
            double *__restrict__ partials, double *gacc, int want_e_, int want_f_) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int gw = blockIdx.x * nw + warp, n_gw = gridDim.x * nw;
    const int F = B.n_feats;
    const bool want_e = want_e_ != 0, want_f = want_f_ != 0;
    unsigned char *mine = smem + (size_t)warp * (size_t)tg.warp_bytes;
    double *acc = GLOBAL_ACC ? gacc + (size_t)gw * 4 * F : (double *)mine;
    unsigned char *scratch = mine + featurize_acc_bytes(F, GLOBAL_ACC);
    TriRec *recs = (TriRec *)scratch;
    const unsigned recs_s = pin(smem_addr(recs));
    typename std::conditional<GLOBAL_ACC, GlobalAcc, SharedAcc>::type acc_rw;
    if constexpr (GLOBAL_ACC) acc_rw.base = acc; else acc_rw.base = pin(smem_addr(acc));
    PairRec *prec = (PairRec *)scratch;
    RoleViews *views = (RoleViews *)(scratch + CHUNK * sizeof(TriRec));

    for (int k = lane; k < 4 * F; k += 32) acc[k] = 0.0;
    __syncwarp();
    constexpr bool LEGS = KP == 3 || KP == 7, CACHED = KP == 7, PLANES = KP >= 4 && KP <= 6;
    constexpr int KC = KP == 4 ? 1 : (KP == 5 ? 2 : 4);
    Tile<((KP == 1 || KP == 2) ? KP : 1)> tile;
    if constexpr (KP > 0 && !PLANES) tile.init(tg, lane);
    int bin0[KC];          // plane path: grid bin of (l0, m, n) for the lane's cells
    if constexpr (PLANES) {
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            const int cell = lane + 32 * kk, ok = cell < tg.ma * tg.na;
            const int mi = ok ? cell / tg.na : 0, ni = ok ? cell % tg.na : 0;
            bin0[kk] = tg.goff + (tg.l0 * tg.dim_m + tg.m0 + mi) * tg.dim_n + tg.n0 + ni;
        }
        double2 *plane = reinterpret_cast<double2 *>(scratch + SP_PLANE);
        for (int k = lane; k < 2 * (tg.ma * tg.na + 1); k += 32) plane[k] = make_double2(0.0, 0.0);
        __syncwarp();
    }

    for (int a = gw; a < f.n; a += n_gw) {
        const int sa = __ldg(f.spec + a);
        const Vec3 pa = real_position(f, a);
        if (lane == 0) acc[4 * sa] += 1.0;      // composition column n_el (composition.py:96-111)
        __syncwarp();

        // ------------------------------------------------ 2-body (bspline.py:810-895)
        two_body_rows(B, f, a, sa, pa, acc, prec, lane);

        // ------------------------------------------------ 3-body (angles.py:17-286)
        if constexpr (PLANES) {
            plane_three_body<KC>(B, f, tg, a, pa, scratch, recs_s, acc_rw.base, bin0, lane, want_e, want_f);
        } else if constexpr (CACHED) {
            legs_three_body_cached(B, f, tg, tg.leg_cache, tg.cache_max3, tg.cache_stride, a, scratch, recs_s,
                                   tile, lane, want_e, want_f);
        } else if constexpr (LEGS) {
            legs_three_body(B, f, tg, a, pa, scratch, recs_s, tile, lane, want_e, want_f);
        } else if (B.n_trios > 0) {
            const int row0 = __ldg(f.off3 + a), n3a = __ldg(f.cnt3 + a);
            // (i) `a` as the centre: every j<k pair of its own list
            const int n_tri = n3a * (n3a - 1) / 2;
            for (int t0 = 0; t0 < n_tri; t0 += CHUNK) {
                const int t = t0 + lane;
                if constexpr (KP > 0) recs[lane].base = 0; else recs[lane].flags = 0;
                if (t < n_tri) {
                    int qj, qk;
                    unrank_pair(t, qj, qk);
                    Triangle T;
                    if (eval_triangle(B, f, pa, sa, __ldg(f.idx3 + row0 + qj), __ldg(f.idx3 + row0 + qk),
                                      0, B.lead3, B.trail3, T))
                        if constexpr (KP > 0) store_tile_record((unsigned char *)(recs + lane), T, tg, 0);
                        else store_record(recs + lane, T, B, 0);
                }
                __syncwarp();
                if constexpr (KP == 1 || KP == 2) tile.accumulate(tg, recs_s, min(CHUNK, n_tri - t0), want_e);
                else scatter_records(B, recs_s, min(CHUNK, n_tri - t0), acc_rw, lane, want_e);
                __syncwarp();
            }
            // (ii) `a` as a neighbour of each centre in its list (force rows only)
            if (want_f) {
                for (int vbase = 0; vbase < n3a; vbase += 32) {
                    const int total = publish_views(B, f, a, vbase, n3a, lane, views);
                    for (int it0 = 0; it0 < total; it0 += CHUNK) {
                        const int it = it0 + lane;
                        if constexpr (KP > 0) recs[lane].base = 0; else recs[lane].flags = 0;
                        if (it < total) {
                            const int v = find_view(views, it);
                            const int ci = views->centre[v], apr = views->a_prime[v];
                            const int mk = __ldg(f.idx3 + __ldg(f.off3 + ci) + (it - views->prefix[v]));
                            if (mk != apr) {
                                Triangle T;
                                const bool first = apr < mk;
                                if (eval_triangle(B, f, real_position(f, ci), __ldg(f.spec + ci),
                                                  first ? apr : mk, first ? mk : apr, first ? 1 : 2,
                                                  B.lead3, B.trail3, T))
                                    if constexpr (KP > 0) store_tile_record((unsigned char *)(recs + lane), T, tg, first ? 1 : 2);
                                    else store_record(recs + lane, T, B, first ? 1 : 2);
                            }
                        }
                        __syncwarp();
                        if constexpr (KP == 1 || KP == 2) tile.accumulate(tg, recs_s, min(CHUNK, total - it0), false);
                        else scatter_records(B, recs_s, min(CHUNK, total - it0), acc_rw, lane, false);
                        __syncwarp();
                    }
                }
            }
        }

        // ------------------------------------------------ rows fx_a, fy_a, fz_a
        if constexpr (KP > 0 && !PLANES) tile.flush(B, tg, acc_rw);
        __syncwarp();
        if (want_f) {
            for (int col = lane; col < F; col += 32) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    __stcs(xf + ((long long)c * f.n + a) * ld + col, acc[4 * col + 1 + c]);   // written once: streaming
                    acc[4 * col + 1 + c] = 0.0;
                }
            }
        }
        __syncwarp();
    }
    if (want_e)
        for (int col = lane; col < F; col += 32) partials[(size_t)gw * F + col] = acc[4 * col];
}

// ---------------------------------------------------------------- cooperative plane kernel
// k_featurize_coop: one BLOCK of W = ceil(cells / 32) warps owns one atom; thread = one
// (m, n) cell of the untrimmed grid and keeps that cell's accumulators for every l in
// REGISTERS ([la][e, fx, fy, fz]).  The plane kernel above spends its time moving
// accumulator quads through shared memory (one read-modify-write per (group, l, cell): the
// LSU data pipe was 81 % busy, half of it bank conflicts); here shared memory only carries
// the sparse leg records and the partner-sum planes:
//   round:  warp w takes leg group g0 + w (groups 0..n3-1: `a` is the centre and the group
//           is its leg to neighbour j; groups n3..2 n3-1: `a` is a neighbour of centre i),
//           evaluates the group's legs and builds ITS plane P[m,n] (, Q[m,n]) as before;
//           __syncthreads; every thread reads its cell from each of the W planes and adds
//           the 4 non-zero l terms of that group to its registers; __syncthreads.
//   atom end: registers are folded into the block's column accumulators (bin_col; bins
//           with l <= m first, then l > m, so the two mirror bins of a column never race)
//           and the three rows are written with all threads.
// Plane layout: two arrays per warp with a row stride `nap` = na rounded up to 4 mod 8
// cells, so the 16 (p, q) products of a partner fall in distinct banks — (P, Qx) and
// (Qy, Qz) as 16-byte entries for the neighbour role, P split over even / odd partners
// as 8-byte entries for the centre role.
constexpr unsigned GL_REC = 160;        // group leg, dense: (V[l], DV[l]) x 8, u[3], {idx, role}
constexpr int CO_LA = 8;                // max untrimmed l extent

struct CoopGeom {
    TileGeom g;
    int warps;              // per block
    int nap;                // plane row stride in cells
    int slots;              // leg records per table
    int off_ltab, off_gleg; // block-shared: legs of the atom's own row; W group legs
    int off_acc2;           // block-shared: second accumulator of the pair columns [col0][4]
    int off_warp, warp_bytes;   // per-warp region: [L records][N records][plane A][plane B][dummy]
    int off_plane;          // of plane A inside the warp region
    int plane_bytes;        // of one plane array
};

// Publish the group's own leg (sparse record `src`) densely for the consumers.
__device__ __forceinline__ void publish_group_leg(unsigned src, unsigned dst, int role, int lane) {
    const int rel = lds32(src + SPL_IDX);
    if (lane < CO_LA) {
        const unsigned r = (unsigned)(lane - rel);
        const bool in = r < 4u;
        const double v = in ? lds64(src + 8u * (in ? r : 0u)) : 0.0;
        const double dv = in ? lds64(src + 32u + 8u * (in ? r : 0u)) : 0.0;
        sts128(dst + 16u * (unsigned)lane, make_double2(v, dv));
    } else if (lane == CO_LA) {
        sts128(dst + 128, lds128(src + 64));                               // ux, uy
    } else if (lane == CO_LA + 1) {
        const long long tag = ((long long)role << 32) | (unsigned)rel;
        sts128(dst + 144, make_double2(lds64(src + 80), __longlong_as_double(tag)));   // uz, {idx, role}
    }
}

// Cached sparse record (global) -> the same record in shared memory; `sign` scales the unit
// vector (the cache stores a pair's leg from the lower to the higher row position).
__device__ __forceinline__ void copy_leg_record(const unsigned char *rec, double sign, unsigned char *out) {
    const double2 *gp = reinterpret_cast<const double2 *>(rec);
    const double2 a0 = __ldg(gp), a1 = __ldg(gp + 1), a2 = __ldg(gp + 2), a3 = __ldg(gp + 3);
    const double2 u01 = __ldg(gp + 4), u2t = __ldg(gp + 5);
    double2 *o = reinterpret_cast<double2 *>(out);
    o[0] = a0; o[1] = a1; o[2] = a2; o[3] = a3;
    o[4] = make_double2(sign * u01.x, sign * u01.y);
    o[5] = make_double2(sign * u2t.x, u2t.y);
}

// One published group added to the thread's registers: the 4 l values IL..IL+3 of the
// group's leg, statically indexed (the switch in add_group picks IL = idx - l0).
template <int LA, int IL, bool CENTRE>
__device__ __forceinline__ void add_group_at(double (&r)[LA][4], unsigned gl, double P, double qx, double qy,
                                             double qz, double ux, double uy, double uz, double half_e) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        constexpr int L0 = IL;
        const int l = L0 + p;
        if (l >= 0 && l < LA) {
            const double2 vd = lds128(gl + 16u * (unsigned)l);     // (B_l, B'_l) of the group's leg
            const double dP = vd.y * P;
            if (CENTRE) {
                r[l][0] = fma(half_e * vd.x, P, r[l][0]);
                r[l][1] = fma(ux, dP, r[l][1]);
                r[l][2] = fma(uy, dP, r[l][2]);
                r[l][3] = fma(uz, dP, r[l][3]);
            } else {
                r[l][1] += vd.x * qx - ux * dP;
                r[l][2] += vd.x * qy - uy * dP;
                r[l][3] += vd.x * qz - uz * dP;
            }
        }
    }
}
template <int LA, bool CENTRE>
__device__ __forceinline__ void add_group(double (&r)[LA][4], int il, unsigned gl, double P, double qx, double qy,
                                          double qz, double ux, double uy, double uz, double half_e) {
#define UF3B_CASE(IL) case IL: if (IL < LA) add_group_at<LA, IL, CENTRE>(r, gl, P, qx, qy, qz, ux, uy, uz, half_e); break;
    switch (il) {
        UF3B_CASE(-3) UF3B_CASE(-2) UF3B_CASE(-1) UF3B_CASE(0) UF3B_CASE(1) UF3B_CASE(2) UF3B_CASE(3)
        UF3B_CASE(4) UF3B_CASE(5) UF3B_CASE(6) UF3B_CASE(7)
        default: break;
    }
#undef UF3B_CASE
}

template <int LA>
__global__ void __launch_bounds__(128, 4)
k_featurize_coop(const BasisTab B, const FrameView f, const CoopGeom cg, double *__restrict__ xf, long long ld,
                 double *__restrict__ partials, int want_e_, int want_f_) {
    extern __shared__ __align__(16) unsigned char smem[];
    const TileGeom &g = cg.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = cg.warps;
    const int F = B.n_feats;
    const bool want_e = want_e_ != 0, want_f = want_f_ != 0;
    double *acc = reinterpret_cast<double *>(smem);                         // [F][e, fx, fy, fz]
    double *acc2 = reinterpret_cast<double *>(smem + cg.off_acc2);          // pair columns, odd passes
    const unsigned smem_s = pin(smem_addr(smem));
    const unsigned ltab_s = smem_s + (unsigned)cg.off_ltab, gleg_s = smem_s + (unsigned)cg.off_gleg;
    unsigned char *mine = smem + cg.off_warp + warp * cg.warp_bytes;
    const unsigned mine_s = smem_s + (unsigned)(cg.off_warp + warp * cg.warp_bytes);
    unsigned char *lm = mine, *nn = mine + cg.slots * SPL_REC;
    const unsigned lm_s = mine_s, nn_s = mine_s + (unsigned)cg.slots * SPL_REC;
    const unsigned pl_a = mine_s + (unsigned)cg.off_plane, pl_b = pl_a + (unsigned)cg.plane_bytes;
    const unsigned dummy_s = pl_b + (unsigned)cg.plane_bytes;
    PairRec *prec = reinterpret_cast<PairRec *>(mine);

    // the thread's cell
    const int n_cells = g.ma * g.na;
    const bool has_cell = tid < n_cells;
    const int mi = has_cell ? tid / g.na : 0, ni = has_cell ? tid % g.na : 0;
    const unsigned cellp = (unsigned)(mi * cg.nap + ni);
    const int lmn = g.dim_m * g.dim_n;
    const int bin_base = g.goff + (g.l0 * g.dim_m + g.m0 + mi) * g.dim_n + g.n0 + ni;
    const unsigned ma = (unsigned)g.ma, na = (unsigned)g.na, nap = (unsigned)cg.nap;
    const int half = lane >> 4, p = (lane >> 2) & 3, q = lane & 3;
    const double half_e = want_e ? 0.5 : 0.0;

    double r[LA][4];
#pragma unroll
    for (int l = 0; l < LA; ++l)
#pragma unroll
        for (int c = 0; c < 4; ++c) r[l][c] = 0.0;
    for (int k = tid; k < 4 * F; k += blockDim.x) acc[k] = 0.0;
    for (int k = tid; k < 4 * g.col0; k += blockDim.x) acc2[k] = 0.0;
    {
        double2 *z = reinterpret_cast<double2 *>(mine + cg.off_plane);
        for (int k = lane; k < (2 * cg.plane_bytes + 32) / 16; k += 32) z[k] = make_double2(0.0, 0.0);
    }
    __syncthreads();

    for (int a = blockIdx.x; a < f.n; a += gridDim.x) {
        const int sa = __ldg(f.spec + a);
        const Vec3 pa = real_position(f, a);
        const int row0 = __ldg(f.off3 + a), n3a = __ldg(f.cnt3 + a);
        // work slots of the atom, dealt to the warps W at a time: slots 0 and 1 = pair rows (even /
        // odd passes, the odd ones into a second small accumulator merged at the fold, so that
        // round 0 is balanced), then the neighbour-role groups, then the centre-role groups.  The
        // legs of a's own row are only read by centre-role groups: the warp of the first group
        // slot loads them in round 0 unless a centre group already falls into round 0.
        const int s2 = W > 1 ? 2 : 1;                 // pair-row slots (even / odd passes of 32 pairs)
        const int c0 = s2 + (want_f ? n3a : 0);       // first centre-role slot
        const int n_slots = c0 + (n3a > 1 ? n3a : 0);
        const bool early = c0 < W;
        const unsigned char *mine_c = g.leg_cache + (size_t)a * g.cache_stride * SPL_REC;
        if (early) {
            if (warp == 0 && lane < n3a)
                copy_leg_record(mine_c + (size_t)lane * SPL_REC, 1.0, smem + cg.off_ltab + lane * SPL_REC);
            __syncthreads();
        }
        for (int g0 = 0; g0 < n_slots; g0 += W) {
            const int slot = g0 + warp;
            if (!early && slot == (W > 1 ? s2 : 0) && lane < n3a)
                copy_leg_record(mine_c + (size_t)lane * SPL_REC, 1.0, smem + cg.off_ltab + lane * SPL_REC);
            if (slot == 0) {
                if (lane == 0) acc[4 * sa] += 1.0;      // composition column n_el (composition.py:96-111)
                __syncwarp();
                two_body_rows(B, f, a, sa, pa, acc, prec, lane, 0, s2);
            } else if (slot == 1 && s2 == 2) {
                two_body_rows(B, f, a, sa, pa, acc2, prec, lane, 1, 2);
            }
            // gs: group index in the old numbering (centre groups 0..n3a-1, neighbour groups n3a..)
            const int gs = slot < s2 ? -1 : (slot < c0 ? n3a + (slot - s2) : (slot < n_slots ? slot - c0 : -1));
            const int n_groups = 2 * n3a;
            int role = 0;                         // 0 none, 1 centre, 2 neighbour
            if (gs >= 0 && gs < n3a) {
                // ---- `a` is the centre, the group is its leg to neighbour j
                const int j = gs;
                if (n3a > 1) {
                    if (lane < n3a && lane != j) {     // leg (j, k = lane): only its values are used here
                        const int lo = j < lane ? j : lane, hi = j < lane ? lane : j;
                        copy_leg_record(mine_c + (size_t)(g.cache_max3 + hi * (hi - 1) / 2 + lo) * SPL_REC, 1.0,
                                        nn + lane * SPL_REC);
                    }
                    __syncwarp();
                    for (int k0 = 0; k0 < n3a; k0 += 2) {     // two partners per pass, one per half-warp
                        const int k = min(k0 + half, n3a - 1);
                        const unsigned lk = ltab_s + (unsigned)k * SPL_REC, ns = nn_s + (unsigned)k * SPL_REC;
                        const unsigned mrel = (unsigned)(lds32(lk + SPL_IDX) + p);
                        const unsigned nrel = (unsigned)(k == j ? SP_DEAD : lds32(ns + SPL_IDX) + q);
                        const bool ok = mrel < ma && nrel < na && k0 + half < n3a;
                        const unsigned ad = ok ? (half ? pl_b : pl_a) + 8u * (mrel * nap + nrel) : dummy_s;
                        sts64(ad, fma(lds64(lk + 8 * p), lds64(ns + 8 * q), lds64(ad)));
                        __syncwarp();
                    }
                    publish_group_leg(ltab_s + (unsigned)j * SPL_REC, gleg_s + (unsigned)warp * GL_REC, 1, lane);
                    role = 1;
                }
            } else if (gs >= n3a && gs < n_groups) {
                // ---- `a` is a neighbour of centre i = entry e of its list
                const int m = __ldg(f.idx3 + row0 + (gs - n3a));
                const int gimg = image_of(f, m);
                const int ci = m - gimg * f.n;
                const int apr = __ldg(f.img_inv + gimg) * f.n + a;
                const int rowi = __ldg(f.off3 + ci), ni_ = __ldg(f.cnt3 + ci);
                const int its = lane < ni_ ? __ldg(f.idx3 + rowi + lane) : -1;
                const unsigned hit = __ballot_sync(FULL, its == apr);
                if (hit) {                        // (a miss: one-ulp asymmetry of the list criterion)
                    const int qa = __ffs(hit) - 1;
                    const unsigned char *theirs = g.leg_cache + (size_t)ci * g.cache_stride * SPL_REC;
                    for (int it0 = 0; it0 < 2 * ni_; it0 += 32) {
                        const int it = it0 + lane;
                        const bool centre_leg = it < ni_;
                        const int k = centre_leg ? it : it - ni_;
                        if (it < 2 * ni_ && (centre_leg || k != qa)) {
                            if (centre_leg) {
                                copy_leg_record(theirs + (size_t)k * SPL_REC, 1.0, lm + k * SPL_REC);
                            } else {              // leg (a', k), unit vector a' -> k
                                const int lo = qa < k ? qa : k, hi = qa < k ? k : qa;
                                copy_leg_record(theirs + (size_t)(g.cache_max3 + hi * (hi - 1) / 2 + lo) * SPL_REC,
                                                qa < k ? 1.0 : -1.0, nn + k * SPL_REC);
                            }
                        }
                    }
                    __syncwarp();
                    const unsigned off_x = 32u * (unsigned)half + 8u * (unsigned)q;      // v[q] | dv[q]
                    const unsigned off_w = 64u + 8u * (unsigned)half;                    // wx | wy
                    const unsigned pl = half ? pl_b : pl_a;
                    for (int k = 0; k < ni_; ++k) {           // lanes 0-15 own (P, Qx), lanes 16-31 (Qy, Qz)
                        if (k == qa) continue;
                        const unsigned lk = lm_s + (unsigned)k * SPL_REC, ns = nn_s + (unsigned)k * SPL_REC;
                        const unsigned mrel = (unsigned)(lds32(lk + SPL_IDX) + p);
                        const unsigned nrel = (unsigned)(lds32(ns + SPL_IDX) + q);
                        const bool ok = mrel < ma && nrel < na;
                        const unsigned ad = ok ? pl + 16u * (mrel * nap + nrel) : dummy_s;
                        const double vm = lds64(lk + 8 * p), x = lds64(ns + off_x), dvn = lds64(ns + 32 + 8 * q);
                        const double wa = lds64(ns + off_w), wz = lds64(ns + 80);
                        double2 c = lds128(ad);
                        c.x = fma(vm * x, half ? wa : 1.0, c.x);       // P += vm vn       | Qy += wy vm dvn
                        c.y = fma(vm * dvn, half ? wz : wa, c.y);      // Qx += wx vm dvn  | Qz += wz vm dvn
                        sts128(ad, c);
                        __syncwarp();
                    }
                    publish_group_leg(lm_s + (unsigned)qa * SPL_REC, gleg_s + (unsigned)warp * GL_REC, 2, lane);
                    role = 2;
                }
            }
            if (role == 0 && lane == 0) sts128(gleg_s + (unsigned)warp * GL_REC + 144, make_double2(0.0, 0.0));
            __syncthreads();

            // ---- consume: every thread adds the W published groups to its cell's registers
            for (int w2 = 0; w2 < W; ++w2) {
                const unsigned gl = gleg_s + (unsigned)w2 * GL_REC;
                const int il = lds32(gl + 152), grole = lds32(gl + 156);
                if (grole == 0) continue;         // block-uniform
                const unsigned wa = smem_s + (unsigned)(cg.off_warp + w2 * cg.warp_bytes + cg.off_plane);
                const unsigned wb = wa + (unsigned)cg.plane_bytes;
                const double2 u01 = lds128(gl + 128);
                const double u2 = lds64(gl + 144);
                if (grole == 1) {
                    double P = 0.0;
                    if (has_cell) {
                        P = lds64(wa + 8u * cellp) + lds64(wb + 8u * cellp);
                        sts64(wa + 8u * cellp, 0.0);
                        sts64(wb + 8u * cellp, 0.0);
                    }
                    add_group<LA, true>(r, il, gl, P, 0.0, 0.0, 0.0, u01.x, u01.y, u2, half_e);
                } else {
                    double2 pq = make_double2(0.0, 0.0), qq = make_double2(0.0, 0.0);
                    if (has_cell) {
                        pq = lds128(wa + 16u * cellp);
                        qq = lds128(wb + 16u * cellp);
                        sts128(wa + 16u * cellp, make_double2(0.0, 0.0));
                        sts128(wb + 16u * cellp, make_double2(0.0, 0.0));
                    }
                    add_group<LA, false>(r, il, gl, pq.x, pq.y, qq.x, qq.y, u01.x, u01.y, u2, half_e);
                }
            }
            __syncthreads();
        }

        // ---- pair columns of the odd passes (3-body columns start at g.col0)
        for (int k = tid; k < 4 * g.col0; k += blockDim.x) { acc[k] += acc2[k]; acc2[k] = 0.0; }
        // ---- fold the registers into the column accumulators: bins with l <= m, then l > m
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
#pragma unroll
            for (int l = 0; l < LA; ++l) {
                if (l < g.la && has_cell && ((g.l0 + l > g.m0 + mi) == (ph == 1))) {
                    const int col = __ldg(B.bin_col + bin_base + l * lmn);
                    if (col >= 0) {
                        double2 *dst = reinterpret_cast<double2 *>(acc + 4 * (size_t)(g.col0 + col));
                        double2 u = dst[0], w = dst[1];
                        u.x += r[l][0]; u.y += r[l][1]; w.x += r[l][2]; w.y += r[l][3];
                        dst[0] = u;
                        dst[1] = w;
                    }
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int l = 0; l < LA; ++l)
#pragma unroll
            for (int c = 0; c < 4; ++c) r[l][c] = 0.0;

        // ---- rows fx_a, fy_a, fz_a
        if (want_f) {
            for (int col = tid; col < F; col += blockDim.x) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    __stcs(xf + ((long long)c * f.n + a) * ld + col, acc[4 * col + 1 + c]);   // written once: streaming
                    acc[4 * col + 1 + c] = 0.0;
                }
            }
        }
        __syncthreads();
    }
    if (want_e)
        for (int col = tid; col < F; col += blockDim.x) partials[(size_t)blockIdx.x * F + col] = acc[4 * col];
}

// Energy row = element counts (composition.py:96-111) + fixed-order sum of the warps'
// partial rows, in two stages so that the sum over thousands of rows is spread over the
// chip: block (x, y) sums rows [y * rows_per_block, ...) of 32 feature columns into
// out[y][col] — warp w takes rows w, w+8, ... with coalesced reads, then the eight per-warp
// sums are added in a fixed order; a second launch folds the ER_SPLIT intermediate rows.
__global__ void __launch_bounds__(256)
k_energy_row(const double *__restrict__ partials, int n_rows, int rows_per_block, int n_feats,
             double *__restrict__ out) {
    __shared__ double red[8][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + lane;
    const int r_begin = blockIdx.y * rows_per_block;
    const int r_end = min(n_rows, r_begin + rows_per_block);
    double s = 0.0;
    if (col < n_feats)      // the element-count columns were accumulated by k_featurize as well
        for (int r = r_begin + warp; r < r_end; r += 8) s += partials[(size_t)r * n_feats + col];
    red[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && col < n_feats) {
        double t = red[0][lane];
#pragma unroll
        for (int w = 1; w < 8; ++w) t += red[w][lane];
        out[(size_t)blockIdx.y * n_feats + col] = t;
    }
}

}  // namespace uf3b

using namespace uf3b;

// partials [n_rows][F] (+ ER_SPLIT scratch rows behind them) -> energy row d_xe
int uf3b::launch_energy_row(double *partials, int n_rows, int F, double *d_xe, cudaStream_t stream) {
    double *mid = partials + (size_t)n_rows * F;
    const int per = (n_rows + ER_SPLIT - 1) / ER_SPLIT;
    UF3B_LAUNCH(k_energy_row, dim3((F + 31) / 32, ER_SPLIT), 256, 0, stream, partials, n_rows, per, F, mid);
    UF3B_LAUNCH(k_energy_row, dim3((F + 31) / 32, 1), 256, 0, stream, mid, ER_SPLIT, ER_SPLIT, F, d_xe);
    return UF3B_OK;
}

// Copies to host buffers (if any), synchronisation and kernel timing shared by the launch paths.
int uf3b::finish_featurize(uf3b_basis *basis, double *x_energy, double *x_forces, int64_t ld, double *d_xe,
                            double *d_xf, int F, int n, bool e_dev, bool f_dev, cudaStream_t stream,
                            cudaEvent_t ev0, cudaEvent_t ev1) {
    bool need_sync = g_timing;
    if (x_forces && !f_dev) {
        if (ld == F)        // contiguous rows: one linear copy
            UF3B_CUDA(cudaMemcpyAsync(x_forces, d_xf, sizeof(double) * F * (size_t)3 * n, cudaMemcpyDeviceToHost, stream));
        else
            UF3B_CUDA(cudaMemcpy2DAsync(x_forces, sizeof(double) * ld, d_xf, sizeof(double) * F,
                                        sizeof(double) * F, (size_t)3 * n, cudaMemcpyDeviceToHost, stream));
        need_sync = true;
    }
    if (x_energy && !e_dev) {
        UF3B_CUDA(cudaMemcpyAsync(x_energy, d_xe, sizeof(double) * F, cudaMemcpyDeviceToHost, stream));
        need_sync = true;
    }
    if (need_sync) UF3B_CUDA(stream_sync(stream));
    if (g_timing) {
        float ms = 0.f;
        UF3B_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        g_last_kernel_ms = ms;
        cudaEventDestroy(ev0);
        cudaEventDestroy(ev1);
    }
    return UF3B_OK;
}

extern "C" int uf3b_featurize(uf3b_basis *basis, const uf3b_nlist *nl, double *x_energy,
                              double *x_forces, int64_t ld, void *stream_) {
    if (!basis || !nl) return fail(UF3B_ERR_INVALID, "null handle");
    DeviceGuard on_device(basis->device);
    // A deferred list build (uf3b_basis_set_deferred_lists) is not waited for when the tiled kernels
    // take the frame with DEVICE outputs: they size their shared memory by the previous frame's longest
    // 3-body row and the caller verifies the build afterwards (nlist_resolve -> UF3B_RETRY).
    uf3b_nlist *nl_rw = const_cast<uf3b_nlist *>(nl);
    bool deferred = nl->pending && basis->deferred_lists && nl->max3_hint >= 2
                    && (!x_energy || is_device_pointer(x_energy)) && (!x_forces || is_device_pointer(x_forces));
    if (!deferred)
        if (int rc = nlist_resolve(nl_rw)) return rc;     // needs the longest row on the host
    cudaStream_t stream = (cudaStream_t)stream_;
    const int F = basis->n_feats;
    const int n = (int)nl->n;
    if (x_forces && ld < F) return fail(UF3B_ERR_INVALID, "ld smaller than n_feats");
    if (nl->c_count < n)
        return fail(UF3B_ERR_STATE, "feature rows need the lists of every centre (built for a centre range)");
    if (!x_energy && !x_forces) return UF3B_OK;
    const bool e_dev = x_energy && is_device_pointer(x_energy);
    const bool f_dev = x_forces && is_device_pointer(x_forces);
    if (n == 0) {
        if (x_energy) {
            if (e_dev) UF3B_CUDA(cudaMemsetAsync(x_energy, 0, sizeof(double) * F, stream));
            else for (int k = 0; k < F; ++k) x_energy[k] = 0.0;
        }
        return UF3B_OK;
    }

    // small 3-body grids of a unary, symmetry-2 basis: the register-tiled kernel (featurize_tiled.cu)
    {
        const int rc = featurize_tiled(basis, nl, x_energy, x_forces, ld, stream, deferred);
        if (rc <= 0) return rc;
        if (deferred)       // another path takes the frame: it needs the verified lists
            if (int rc2 = nlist_resolve(nl_rw)) return rc2;
    }
    // several species, symmetry-1 trios, long rows: the general leg-grouped kernel (featurize_multi.cu)
    if (!basis->no_tile) {
        const int rc = featurize_multi(basis, nl, x_energy, x_forces, ld, stream);
        if (rc <= 0) return rc;
    }
    // launch shape: as many warps per SM as shared memory and registers allow, grid sized to
    // the SM count
    int dev = 0, smem_max = 0, smem_sm = 0;
    UF3B_CUDA(cudaGetDevice(&dev));
    UF3B_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    UF3B_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    // accumulators in shared memory while at least two warps fit a block, else in global memory
    const bool global_acc = 2 * featurize_warp_bytes(F, false) > (size_t)smem_max;
    // register-tile path: unary basis, unit folding weights, small untrimmed 3-body grid
    TileGeom tg = {};
    int kp = 0;
    if (!global_acc && basis->tab.n_trios == 1 && basis->tab.unit_weights && !basis->no_tile) {
        const int lead = basis->tab.lead3, trail = basis->tab.trail3;
        const int L = basis->h_trio_dims[0], M = basis->h_trio_dims[1], N = basis->h_trio_dims[2];
        tg.l0 = tg.m0 = tg.n0 = lead;
        tg.la = L - lead - trail; tg.ma = M - lead - trail; tg.na = N - lead - trail;
        tg.dim_m = M; tg.dim_n = N;
        tg.goff = 0; tg.col0 = basis->h_trio_col[0]; tg.sym = basis->h_trio_sym[0];
        if (tg.la >= 1 && tg.la <= RT_LA && tg.ma >= 1 && tg.na >= 1 && tg.ma * tg.na <= 64
            && tg.ma + tg.na <= 12 && tg.sym >= 1 && tg.sym <= 3)
            kp = tg.ma * tg.na <= 32 ? 1 : 2;
        // leg-grouped path: l and m legs interchangeable, every 3-body row fits one warp pass
        if (kp == 1 && tg.sym >= 2 && tg.la == tg.ma && tg.na <= 12 && nl->max3 <= 32 && !getenv("UF3B_NO_LEGS"))
            kp = 3;
        // plane path: the same factorisation for larger grids (symmetry 2, up to 128 (m, n) cells)
        const bool planes_ok = tg.sym == 2 && tg.la == tg.ma && tg.la >= 1 && tg.na >= 1
                               && tg.ma * tg.na <= PL_MAX_CELLS && nl->max3 <= 32 && !getenv("UF3B_NO_LEGS")
                               && !getenv("UF3B_NO_PLANES");
        if (planes_ok && (kp == 0 || getenv("UF3B_PLANES")))
            kp = tg.ma * tg.na <= 32 ? 4 : (tg.ma * tg.na <= 64 ? 5 : 6);
    }
    // cooperative plane kernel: one block per atom, thread = (m, n) cell, registers hold the l axis
    // leg cache for the leg-grouped path: every distinct leg evaluated once by k_leg_cache
    if (kp == 3 && x_forces && !getenv("UF3B_NO_LEG_CACHE")) {
        const int max3 = std::max(nl->max3, 1);
        const long long stride = max3 + (long long)max3 * (max3 - 1) / 2;
        const size_t bytes = (size_t)n * (size_t)stride * SPL_REC;
        if (bytes <= ((size_t)2 << 30)) {
            UF3B_CUDA(basis->leg_cache.reserve(bytes));
            tg.leg_cache = basis->leg_cache.p;
            tg.cache_max3 = max3;
            tg.cache_stride = (int)stride;
            kp = 7;
        }
    }
    bool coop = kp >= 4 && kp <= 6 && tg.la <= CO_LA && !getenv("UF3B_NO_COOP");
    if (coop) {         // the cooperative kernel reads every leg from the leg cache
        const int max3 = std::max(nl->max3, 1);
        const long long stride = max3 + (long long)max3 * (max3 - 1) / 2;
        const size_t bytes = (size_t)n * (size_t)stride * SPL_REC;
        if (bytes <= ((size_t)2 << 30)) {
            UF3B_CUDA(basis->leg_cache.reserve(bytes));
            tg.leg_cache = basis->leg_cache.p;
            tg.cache_max3 = max3;
            tg.cache_stride = (int)stride;
        } else {
            coop = false;
        }
    }
    if (coop) {
        CoopGeom cg = {};
        cg.g = tg;
        const int n_cells = tg.ma * tg.na;
        cg.warps = (n_cells + 31) / 32;
        if (cg.warps == 3) cg.warps = 4;
        cg.nap = tg.na + ((4 - tg.na % 8) + 8) % 8;
        cg.slots = std::max(16, std::min(32, (nl->max3 + 3) & ~3));
        cg.plane_bytes = tg.ma * cg.nap * 16;
        cg.off_ltab = (int)featurize_acc_bytes(F, false);
        cg.off_gleg = cg.off_ltab + cg.slots * (int)SPL_REC;
        cg.off_acc2 = cg.off_gleg + cg.warps * (int)GL_REC;
        cg.off_warp = cg.off_acc2 + 32 * tg.col0;
        cg.off_plane = 2 * cg.slots * (int)SPL_REC;
        cg.warp_bytes = cg.off_plane + 2 * cg.plane_bytes + 32;
        const size_t smem_c = (size_t)cg.off_warp + (size_t)cg.warps * cg.warp_bytes;
        auto kc = tg.la <= 4 ? k_featurize_coop<4> : k_featurize_coop<8>;
        if (smem_c <= (size_t)smem_max) {
            UF3B_CUDA(ensure_dynamic_smem((const void *)kc, smem_c));
            int per_sm = 1;
            UF3B_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kc, cg.warps * 32, smem_c));
            if (per_sm < 1) per_sm = 1;
            int grid = std::min(std::max(1, sm_count() * per_sm / std::max(1, basis->frames_in_flight)), n);
            double *d_xf = x_forces;
            long long d_ld = ld;
            if (x_forces && !f_dev) {
                UF3B_CUDA(basis->stage.reserve((size_t)3 * n * F));
                d_xf = basis->stage.p;
                d_ld = F;
            }
            double *d_xe = x_energy;
            if (x_energy) {
                UF3B_CUDA(basis->partials.reserve((size_t)(grid + ER_SPLIT) * F));
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
            UF3B_LAUNCH(k_leg_cache, n, 128, 0, stream, basis->tab, view, tg,
                        tg.cache_max3, tg.cache_stride, basis->leg_cache.p);
            UF3B_LAUNCH(kc, grid, cg.warps * 32, smem_c, stream, basis->tab, view, cg, d_xf, d_ld,
                        basis->partials.p, x_energy ? 1 : 0, x_forces ? 1 : 0);
            if (g_timing) UF3B_CUDA(cudaEventRecord(ev1, stream));
            if (x_energy)
                if (int rc = launch_energy_row(basis->partials.p, grid, F, d_xe, stream)) return rc;
            return finish_featurize(basis, x_energy, x_forces, ld, d_xe, d_xf, F, n, e_dev, f_dev, stream, ev0, ev1);
        }
    }
    const size_t scratch_bytes = (kp >= 4 && kp <= 6) ? plane_scratch_bytes(tg.ma * tg.na) : ((WARP_SCRATCH + 15) & ~size_t(15));
    const size_t per_warp = featurize_acc_bytes(F, global_acc) + scratch_bytes;
    tg.warp_bytes = (int)per_warp;
    // warps per block: the count that keeps most warps resident (128 registers per thread
    // allow 16 warps per SM; each block also pays 1 KB of reserved shared memory)
    int warps = 1, best = 0;
    for (int w = global_acc ? 4 : 8; w >= 1; --w) {
        const size_t blk = (size_t)w * per_warp;
        if (blk > (size_t)smem_max) continue;
        const int resident = std::min((int)((size_t)smem_sm / (blk + 1024)), 16 / w) * w;
        if (resident > best) { best = resident; warps = w; }
    }
    const size_t smem = (size_t)warps * per_warp;
    auto kernel = k_featurize<false, 0>;
    switch (global_acc ? -1 : kp) {
        case -1: kernel = k_featurize<true, 0>; break;
        case 1: kernel = k_featurize<false, 1>; break;
        case 2: kernel = k_featurize<false, 2>; break;
        case 3: kernel = k_featurize<false, 3>; break;
        case 4: kernel = k_featurize<false, 4>; break;
        case 5: kernel = k_featurize<false, 5>; break;
        case 6: kernel = k_featurize<false, 6>; break;
        case 7: kernel = k_featurize<false, 7>; break;
        default: break;
    }
    UF3B_CUDA(ensure_dynamic_smem((const void *)kernel, smem));
    int per_sm = 1;
    UF3B_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, warps * 32, smem));
    if (per_sm < 1) per_sm = 1;
    // `waves` resident grids: 1 = persistent blocks; more lets blocks retire during the launch,
    // so kernels of another stream (the next frame's list build) can be scheduled in between
    static const int waves = getenv("UF3B_GRID_WAVES") ? std::max(1, atoi(getenv("UF3B_GRID_WAVES"))) : 1;
    // frames in flight (uf3b_basis_set_frames_in_flight): a launch takes only 1/k of the resident
    // blocks, so that k frames on different streams share every SM — the tail of one frame's
    // kernel and the next frame's list build overlap instead of leaving SMs idle
    int grid = std::max(1, sm_count() * per_sm * waves / std::max(1, basis->frames_in_flight));
    const int need = (n + warps - 1) / warps;
    if (grid > need) grid = need;
    const int n_gw = grid * warps;

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
    if (global_acc) UF3B_CUDA(basis->gacc.reserve((size_t)n_gw * 4 * F));
    if (kp == 7)
        UF3B_LAUNCH(k_leg_cache, n, 128, 0, stream, basis->tab, view, tg,
                    tg.cache_max3, tg.cache_stride, basis->leg_cache.p);
    UF3B_LAUNCH(kernel, grid, warps * 32, smem, stream, basis->tab, view, tg, d_xf, d_ld,
                basis->partials.p, basis->gacc.p, x_energy ? 1 : 0, x_forces ? 1 : 0);
    if (g_timing) UF3B_CUDA(cudaEventRecord(ev1, stream));
    if (x_energy)
        if (int rc = launch_energy_row(basis->partials.p, n_gw, F, d_xe, stream)) return rc;
    return finish_featurize(basis, x_energy, x_forces, ld, d_xe, d_xf, F, n, e_dev, f_dev, stream, ev0, ev1);
}
