// spline.cuh — cubic B-spline evaluation from per-interval polynomial pieces.
//
// Reference semantics reproduced (representation/bspline.py:950-974, :791-849):
//   idx = searchsorted(knots, r, 'left') - 4  -> first of the four non-zero basis
//   functions; r on a knot belongs to the interval on its LEFT; the point contributes
//   nothing unless 0 <= idx <= n_basis - 4, i.e. knots[3] < r <= knots[n_knots-4].
// scipy evaluates each basis element by de Boor recursion; here the recursion is run
// once per knot interval on the host (long double, polynomial arithmetic) and the
// kernels evaluate cubics in the local variable u = r - t[i] by Horner's rule.
#pragma once
#include <vector>

#include "common.cuh"

namespace uf3b {

#if defined(__CUDACC__)
#define UF3B_HD __host__ __device__ __forceinline__
#else
#define UF3B_HD inline
#endif

// Interval i with t[i] < r <= t[i+1], 3 <= i <= nk-5, or -1.
// `scale` = (nk - 7) / (t[nk-4] - t[3]) (intervals per unit length, precomputed).
UF3B_HD int find_interval(const double *t, int nk, double scale, double r) {
    const double t_first = t[3], t_last = t[nk - 4];
    if (!(r > t_first) || !(r <= t_last)) return -1;
    // uniform-spacing guess, verified against the real knots
    int i = 3 + (int)((r - t_first) * scale);
    if (i > nk - 5) i = nk - 5;
    if (t[i] < r && r <= t[i + 1]) return i;
    int lo = 3, hi = nk - 4;    // invariant: t[lo] < r <= t[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (t[mid] < r) lo = mid; else hi = mid;
    }
    return lo;
}

// Values and first derivatives of basis functions i-3..i at r (piece = poly + PS*(i-3)).
// Pieces are 16-byte aligned (16 doubles each, PS doubles apart: 16 in the global tables; tables
// staged in shared memory use PS = 18, because with the natural 128-byte stride every lane whose
// distance falls into a different knot interval hits the same banks).
// NC: read through the non-coherent global path (__ldg); false for tables staged in shared memory.
template <bool NC = true>
UF3B_HD void eval_piece(const double *piece, double u, double v[4], double dv[4]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#if defined(__CUDA_ARCH__)
        const double2 *p2 = reinterpret_cast<const double2 *>(piece) + 2 * q;
        const double2 lo = NC ? __ldg(p2) : p2[0];
        const double2 hi = NC ? __ldg(p2 + 1) : p2[1];
        const double c0 = lo.x, c1 = lo.y, c2 = hi.x, c3 = hi.y;
#else
        const double c0 = piece[4 * q + 0], c1 = piece[4 * q + 1];
        const double c2 = piece[4 * q + 2], c3 = piece[4 * q + 3];
#endif
        v[q] = ((c3 * u + c2) * u + c1) * u + c0;
        dv[q] = (3.0 * c3 * u + 2.0 * c2) * u + c1;
    }
}

// Full leg evaluation with trims (angles.py:554-565, bspline.py:840): returns the
// first basis index or -1; basis indices outside [n_lead, n_basis - n_trail) give 0.
template <bool NC = true, int PS = 16>
UF3B_HD int eval_leg(const double *t, int nk, double scale, const double *poly, double r, int n_lead,
                     int n_trail, double v[4], double dv[4]) {
    const int i = find_interval(t, nk, scale, r);
    if (i < 0) return -1;
    eval_piece<NC>(poly + PS * (i - 3), r - t[i], v, dv);
    const int idx = i - 3, nb = nk - 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int bi = idx + q;
        if (bi < n_lead || bi >= nb - n_trail) { v[q] = 0.0; dv[q] = 0.0; }
    }
    return idx;
}

inline double knot_scale(const double *t, int nk) {
    const double span = t[nk - 4] - t[3];
    return span > 0.0 ? (double)(nk - 7) / span : 0.0;
}

// Host: polynomial pieces of a knot vector (Cox-de Boor on polynomials in u).
inline void build_pieces(const double *t, int nk, std::vector<double> &out) {
    const int n_piece = nk - 7;
    for (int piece = 0; piece < n_piece; ++piece) {
        const int i = piece + 3;
        long double cur[5][4] = {};   // cur[s] = N_{i-3+s, j}(u); cur[4] stays 0
        cur[3][0] = 1.0L;
        for (int j = 1; j <= 3; ++j) {
            long double nxt[5][4] = {};
            for (int s = 0; s < 4; ++s) {
                const int r = i - 3 + s;
                if (r < i - j || r > i) continue;
                const long double d1 = (long double)t[r + j] - (long double)t[r];
                if (d1 != 0.0L) {   // ((u + (t_i - t_r)) / d1) * cur[s]
                    const long double a0 = ((long double)t[i] - (long double)t[r]) / d1;
                    const long double a1 = 1.0L / d1;
                    for (int d = 0; d < 4; ++d) {
                        nxt[s][d] += a0 * cur[s][d];
                        if (d > 0) nxt[s][d] += a1 * cur[s][d - 1];
                    }
                }
                const long double d2 = (long double)t[r + j + 1] - (long double)t[r + 1];
                if (d2 != 0.0L) {   // (((t_{r+j+1} - t_i) - u) / d2) * cur[s+1]
                    const long double b0 = ((long double)t[r + j + 1] - (long double)t[i]) / d2;
                    const long double b1 = -1.0L / d2;
                    for (int d = 0; d < 4; ++d) {
                        nxt[s][d] += b0 * cur[s + 1][d];
                        if (d > 0) nxt[s][d] += b1 * cur[s + 1][d - 1];
                    }
                }
            }
            for (int s = 0; s < 5; ++s)
                for (int d = 0; d < 4; ++d) cur[s][d] = nxt[s][d];
        }
        for (int s = 0; s < 4; ++s)
            for (int d = 0; d < 4; ++d) out.push_back((double)cur[s][d]);
    }
}

}  // namespace uf3b
