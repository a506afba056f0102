// triangle.cuh — one 3-body term seen from the atom that accumulates it.
//
// The reference walks every centre of the ghost supercell and scatters each triangle's
// derivative into dense per-atom grids (representation/angles.py:142-286,424-514).  The
// kernels here use an owner-computes form instead: every real atom `a` visits the
// triangles it takes part in — as the centre, or as one of the two neighbours of a
// centre i taken from a's own 3-body list — and adds only its own share, so rows and
// forces are written without atomics.  A triangle is always evaluated in the frame of
// its REAL centre i (neighbours = supercell indices from i's sorted list), which keeps
// the j<k index rule (angles.py:460-476), the atomic-number reordering (:482-488), the
// inclusive leg filter (:502-508) and every distance bit-identical to the reference's
// real-centre enumeration; ghost-centred triangles of the reference are periodic
// images of these.
#pragma once
#include "geom.cuh"
#include "spline.cuh"

namespace uf3b {

struct Triangle {
    double v[3][4], dv[3][4];   // legs l=(i,j), m=(i,k), n=(j,k): 4 values / derivatives
    double A[3], B[3], C[3];    // d(r_l)/dR_a, d(r_m)/dR_a, d(r_n)/dR_a with the sign of
                                // the accumulated quantity (-dE/dR_a)
    int trio;                   // interaction index  centre*n_pairs + pair(j,k)
    int il, im, in;             // first basis index per leg
    int dim_m, dim_n;           // grid extents of legs m and n
    int lo[3], cnt[3];          // untrimmed part [lo, lo+cnt) of each leg's 4 basis functions
    double ujk[3];              // unit vector j -> k (after the atomic-number reordering)
    double r[3];                // leg lengths r_ij, r_ik, r_jk (after the reordering)
    int atom_j, atom_k;         // parent atoms of j and k (after the reordering)
};

// role: 0 = `a` is the centre, 1 = `a` is the neighbour with supercell index mj,
// 2 = `a` is the neighbour with index mk (mj < mk).  Returns false when the triangle
// contributes nothing.
// Same with the two neighbours given by position, parent atom and species (mj < mk by
// supercell index, as in the centre's sorted list).
template <bool NC = true, int PS = 16>
__device__ __forceinline__ bool eval_triangle_at(const BasisTab &B, const Vec3 &pc, int sc, Vec3 pj, int aj,
                                                 int sj, Vec3 pk, int ak, int sk, int role, int n_lead,
                                                 int n_trail, Triangle &T) {
    double dij = dist_rn(pc, pj), dik = dist_rn(pc, pk);
    if (sj > sk) {                              // angles.py:482-488 (stable by atomic number)
        Vec3 tp = pj; pj = pk; pk = tp;
        double td = dij; dij = dik; dik = td;
        int ts = sj; sj = sk; sk = ts;
        ts = aj; aj = ak; ak = ts;
        role = role == 1 ? 2 : (role == 2 ? 1 : 0);
    }
    const int t = sc * B.n_pairs + pair_index(B.ne, sj, sk);
    const int nkl = __ldg(B.trio_nk + 3 * t), nkm = __ldg(B.trio_nk + 3 * t + 1);
    const int nkn = __ldg(B.trio_nk + 3 * t + 2);
    const double *tl = B.knots3 + __ldg(B.trio_koff + 3 * t);
    const double *tm = B.knots3 + __ldg(B.trio_koff + 3 * t + 1);
    const double *tn = B.knots3 + __ldg(B.trio_koff + 3 * t + 2);
    const double djk = dist_rn(pj, pk);
    if (!(dij >= tl[0] && dij <= tl[nkl - 1])) return false;      // angles.py:502-508
    if (!(dik >= tm[0] && dik <= tm[nkm - 1])) return false;
    if (!(djk >= tn[0] && djk <= tn[nkn - 1])) return false;
    T.il = eval_leg<NC, PS>(tl, nkl, __ldg(B.trio_scale + 3 * t), B.poly3 + __ldg(B.trio_poff + 3 * t) / 16 * PS, dij, n_lead, n_trail, T.v[0], T.dv[0]);
    T.im = eval_leg<NC, PS>(tm, nkm, __ldg(B.trio_scale + 3 * t + 1), B.poly3 + __ldg(B.trio_poff + 3 * t + 1) / 16 * PS, dik, n_lead, n_trail, T.v[1], T.dv[1]);
    T.in = eval_leg<NC, PS>(tn, nkn, __ldg(B.trio_scale + 3 * t + 2), B.poly3 + __ldg(B.trio_poff + 3 * t + 2) / 16 * PS, djk, n_lead, n_trail, T.v[2], T.dv[2]);
    if (T.il < 0 || T.im < 0 || T.in < 0) return false;           // r exactly on the first knot
    T.trio = t;
    T.dim_m = nkm - 4;
    T.dim_n = nkn - 4;
    {   // basis indices kept by the trims: [n_lead, n_basis - n_trail)  (angles.py:554)
        const int first[3] = {T.il, T.im, T.in}, nb[3] = {nkl - 4, nkm - 4, nkn - 4};
#pragma unroll
        for (int leg = 0; leg < 3; ++leg) {
            const int lo = max(0, n_lead - first[leg]), hi = min(4, nb[leg] - n_trail - first[leg]);
            T.lo[leg] = lo;
            T.cnt[leg] = max(0, hi - lo);
        }
        if (T.cnt[0] == 0 || T.cnt[1] == 0 || T.cnt[2] == 0) return false;
    }
    // direction cosines (distances.py:354-363): u_ab = (x_b - x_a) / r_ab
    const double il_ = fast_rcp(dij), im_ = fast_rcp(dik), in_ = fast_rcp(djk);
    const double uij[3] = {(pj.x - pc.x) * il_, (pj.y - pc.y) * il_, (pj.z - pc.z) * il_};
    const double uik[3] = {(pk.x - pc.x) * im_, (pk.y - pc.y) * im_, (pk.z - pc.z) * im_};
    const double ujk[3] = {(pk.x - pj.x) * in_, (pk.y - pj.y) * in_, (pk.z - pj.z) * in_};
    T.atom_j = aj;
    T.atom_k = ak;
    T.r[0] = dij; T.r[1] = dik; T.r[2] = djk;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        T.ujk[c] = ujk[c];
        // angles.py:282-285 then force_grids -= grids (:229-231), per participating atom
        T.A[c] = role == 0 ? uij[c] : (role == 1 ? -uij[c] : 0.0);
        T.B[c] = role == 0 ? uik[c] : (role == 2 ? -uik[c] : 0.0);
        T.C[c] = role == 0 ? 0.0 : (role == 1 ? ujk[c] : -ujk[c]);
    }
    return true;
}

template <bool NC = true, int PS = 16>
__device__ __forceinline__ bool eval_triangle(const BasisTab &B, const FrameView &f, const Vec3 &pc,
                                              int sc, int mj, int mk, int role, int n_lead,
                                              int n_trail, Triangle &T) {
    int aj, ak;
    const Vec3 pj = super_position(f, mj, aj), pk = super_position(f, mk, ak);
    return eval_triangle_at<NC, PS>(B, pc, sc, pj, aj, __ldg(f.spec + aj), pk, ak, __ldg(f.spec + ak), role, n_lead,
                                n_trail, T);
}

// t in [0, n(n-1)/2)  ->  (qj < qk), enumerated qk-major.
__device__ __forceinline__ void unrank_pair(int t, int &qj, int &qk) {
    int k = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)t)) * 0.5f);
    while (k * (k - 1) / 2 > t) --k;
    while ((k + 1) * k / 2 <= t) ++k;
    qk = k;
    qj = t - k * (k - 1) / 2;
}

// The neighbour-role work of atom `a`, 32 list entries at a time: entry e of a's 3-body
// list names a centre i (image g); `a` appears in i's own list as image inv(g).  Lane l
// publishes the view of entry vbase+l; returns the number of (view, k) items.
struct RoleViews {
    int centre[32];
    int a_prime[32];     // supercell index of `a` as seen from the centre
    int prefix[33];      // exclusive prefix of the centres' row lengths
};

__device__ __forceinline__ int publish_views(const BasisTab &B, const FrameView &f, int a, int vbase,
                                             int n3a, int lane, RoleViews *vw) {
    const int e = vbase + lane;
    int cnt = 0, ci = 0, apr = 0;
    if (e < n3a) {
        const int m = __ldg(f.idx3 + __ldg(f.off3 + a) + e);
        const int g = image_of(f, m);
        ci = m - g * f.n;
        apr = __ldg(f.img_inv + g) * f.n + a;
        int dummy;
        const Vec3 pi = real_position(f, ci), pa = super_position(f, apr, dummy);
        const double d = dist_rn(pi, pa);      // same expression the list kernel evaluated
        if (d > B.r3min && d <= B.r3max) cnt = __ldg(f.cnt3 + ci);
    }
    int inc = cnt;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const int up = __shfl_up_sync(FULL, inc, s);
        if (lane >= s) inc += up;
    }
    vw->centre[lane] = ci;
    vw->a_prime[lane] = apr;
    vw->prefix[lane] = inc - cnt;
    const int total = __shfl_sync(FULL, inc, 31);
    if (lane == 31) vw->prefix[32] = total;
    __syncwarp();
    return total;
}

// item -> (view, position in the centre's row); prefix is non-decreasing.
__device__ __forceinline__ int find_view(const RoleViews *vw, int item) {
    int lo = 0, hi = 32;         // prefix[lo] <= item < prefix[hi]
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int mid = (lo + hi) >> 1;
        if (vw->prefix[mid] <= item) lo = mid; else hi = mid;
    }
    return lo;
}

}  // namespace uf3b
