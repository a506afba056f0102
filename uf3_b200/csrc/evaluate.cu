// evaluate.cu — Kernel B (inference path): energy and forces from fitted coefficients.
//
// Replaces UFCalculator._energy_1b/_energy_2b/_energy_3b and _forces_2b/_forces_3b
// (forcefield/calculator.py:183-343), which call ndsplines.NDSpline on dense distance
// and direction-cosine arrays.  One warp owns one real atom; lanes walk its pairs and
// the triangles it takes part in (triangle.cuh), contract the 4 (pair) or 4x4x4
// (triplet) non-zero basis products against the coefficient vector / decompressed
// coefficient grid, and the atom's force is reduced across the warp with shuffles.
// As in the reference, no trimming is applied at evaluation time: trimmed basis
// functions carry zero coefficients (calculator.py:207,286,565-571).
//
// Two force schemes share the code.  NEWTON (default): every triangle is evaluated once,
// by its centre, and the reactions on the two neighbours are added to their parent atoms
// with float64 atomics (red.global.add.f64) — a third of the work; the summation order of
// those atomics varies from run to run at the 1e-16 level.  Owner-computes
// (UF3B_DETERMINISTIC_FORCES=1): each atom also visits the triangles of the centres in
// its list and adds only its own share — no atomics, bit-reproducible, 3x the triangles.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "geom.cuh"
#include "spline.cuh"
#include "triangle.cuh"

namespace uf3b {

// value and the three leg-partials of  sum_pqr C[il+p, im+q, in+r] Bl_p Bm_q Bn_r
template <class Ptr>
__device__ __forceinline__ void contract(Ptr grid, const Triangle &T,
                                         double &val, double &gl, double &gm, double &gn) {
    val = gl = gm = gn = 0.0;
    const int mn = T.dim_m * T.dim_n;
    const double *base = grid + (T.il * T.dim_m + T.im) * T.dim_n + T.in;   // global or shared
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        double u0 = 0.0, u1 = 0.0, u2 = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double *row = base + p * mn + q * T.dim_n;
            double t0 = 0.0, t1 = 0.0;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const double c = row[r];
                t0 += c * T.v[2][r];
                t1 += c * T.dv[2][r];
            }
            u0 += t0 * T.v[1][q];
            u1 += t0 * T.dv[1][q];
            u2 += t1 * T.v[1][q];
        }
        val += u0 * T.v[0][p];
        gl += u0 * T.dv[0][p];
        gm += u1 * T.v[0][p];
        gn += u2 * T.v[0][p];
    }
}

constexpr int EV_WARPS = 4;

// grid_in_smem: the decompressed coefficient grids (n_grid doubles) are staged in shared
// memory once per block.
// VIRIAL: also accumulate W = sum over every leg of every energy term of
// (dE/dr_leg) r_leg u (x) u  (= dE/d(strain) under a homogeneous deformation; stress = W / V),
// the analytic form of what the reference obtains by finite differences over strained
// cells (calculator.py:399-404).  Six per-lane accumulators, written next to the per-warp
// energies: e_partials[n_gw * (1 + c) + gw], c = xx, yy, zz, yz, xz, xy.
// Table sizes (doubles) for staging the spline tables in shared memory; 0 = leave in global.
struct EvalStage { int knots2, poly2, knots3, poly3; };

// One neighbour of the centre in the per-warp table (leg table path, see k_energy_forces).
// x y z {atom | species} | (B, dB) x 4 of the leg (centre, neighbour) | r {first index, lo | cnt << 8} | pad:
// 144 bytes, an odd multiple of 16, so that lanes reading the same field of different records spread over
// the banks (with 128 the table reads cost 4-5 wavefronts too many each)
constexpr unsigned NB_REC = 144;
constexpr int PS_SMEM = 18;         // doubles between polynomial pieces staged in shared memory (spline.cuh)

// PADDED: the spline tables are staged in shared memory, pieces PS_SMEM doubles apart.
// leg_table (unary basis whose two centre legs share their knots, NEWTON scheme): every leg
// (centre, neighbour) is evaluated ONCE per centre into the warp's neighbour table — the per-triangle
// form evaluated it once per triangle, 13 times in bulk W — with the trims applied, and the 4 x 4 x 4
// contraction only walks the basis functions the trims keep (their coefficients are the only
// non-zero ones: decompress_3B scatters into the untrimmed bins, bspline.py:693-719), 2 x 2 x 4 of
// the 64 terms on average for the W model.
template <bool NEWTON, bool VIRIAL, bool PADDED>
__global__ void __launch_bounds__(EV_WARPS * 32, 4)
k_energy_forces(const BasisTab B_, const FrameView f, double *forces,
                double *__restrict__ e_partials, int want_e_, int want_f_, int n_grid, int grid_in_smem,
                const EvalStage st, int leg_table) {
    constexpr int PS = PADDED ? PS_SMEM : 16;
    __shared__ RoleViews s_views[EV_WARPS];
    __shared__ __align__(16) unsigned char s_nbr_raw[EV_WARPS * 32 * NB_REC];
    extern __shared__ __align__(16) double s_grid[];
    // knots and polynomial pieces are read ~14 times per leg: staged in shared memory next to
    // the coefficient grids (the capture of the global-table version showed long-scoreboard
    // stalls of 7.7 cycles per issued instruction; most of the rest is the position gathers)
    BasisTab B = B_;
    {
        double *dst = s_grid + (grid_in_smem ? ((n_grid + 1) & ~1) : 0);
        if (st.knots2 > 0) {
            for (int k = threadIdx.x; k < st.knots2; k += blockDim.x) dst[k] = B_.knots2[k];
            B.knots2 = dst; dst += (st.knots2 + 1) & ~1;
            for (int k = threadIdx.x; k < st.poly2; k += blockDim.x) dst[(k >> 4) * PS + (k & 15)] = B_.poly2[k];
            B.poly2 = dst; dst += (st.poly2 / 16) * PS;
        }
        if (st.knots3 > 0) {
            for (int k = threadIdx.x; k < st.knots3; k += blockDim.x) dst[k] = B_.knots3[k];
            B.knots3 = dst; dst += (st.knots3 + 1) & ~1;
            for (int k = threadIdx.x; k < st.poly3; k += blockDim.x) dst[(k >> 4) * PS + (k & 15)] = B_.poly3[k];
            B.poly3 = dst; dst += (st.poly3 / 16) * PS;
        }
    }
    if (grid_in_smem)
        for (int k = threadIdx.x; k < n_grid; k += blockDim.x) s_grid[k] = B.c_grid[k];
    __syncthreads();
    const double *c_grid = grid_in_smem ? s_grid : B.c_grid;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * EV_WARPS + warp, n_gw = gridDim.x * EV_WARPS;
    const bool want_e = want_e_ != 0, want_f = want_f_ != 0;
    RoleViews *views = s_views + warp;
    double e_acc = 0.0;
    double w[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    auto add_virial = [&](double k, double x, double y, double z) {   // k = (dE/dr) r, u = (x, y, z)
        const double kx = k * x, ky = k * y;
        w[0] += kx * x; w[1] += ky * y; w[2] += k * z * z;
        w[3] += ky * z; w[4] += kx * z; w[5] += kx * y;
    };

    // (centres dealt out through an atomic counter instead of this fixed stride were measured: 2 % on
    // the 12 500-centre share of one of eight ranks, and the energy sum lost its run-to-run bit
    // reproducibility — not kept)
    for (int a = f.c_first + gw; a < f.c_first + f.c_count; a += n_gw) {
        const int sa = __ldg(f.spec + a);
        const Vec3 pa = real_position(f, a);
        double fx = 0.0, fy = 0.0, fz = 0.0;
        if (lane == 0) e_acc += __ldg(B.coeff + sa);                    // calculator.py:183-189

        // ---- 2-body: E += S(r) per ordered pair; F_a = 2 sum_j S'(r_aj) (x_j - x_a)/r_aj
        const int r0 = __ldg(f.off2 + a), r1 = r0 + __ldg(f.cnt2 + a);
        for (int e = r0 + lane; e < r1; e += 32) {
            int aj;
            const Vec3 pj = super_position(f, __ldg(f.idx2 + e), aj);
            const double d = dist_rn(pa, pj);
            const int pr = pair_index(B.ne, sa, __ldg(f.spec + aj));
            double v[4], dv[4];
            const int idx = eval_leg<false, PS>(B.knots2 + __ldg(B.pair_koff + pr), __ldg(B.pair_nk + pr), __ldg(B.pair_scale + pr),
                                                B.poly2 + __ldg(B.pair_poff + pr) / 16 * PS, d, 0, 0, v, dv);
            if (idx < 0) continue;
            const double *c = B.coeff + __ldg(B.pair_col + pr) + idx;
            double s = 0.0, ds = 0.0;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const double cr = __ldg(c + r);
                s += cr * v[r];
                ds += cr * dv[r];
            }
            e_acc += s;
            const double k = 2.0 * ds / d;
            fx += k * (pj.x - pa.x);
            fy += k * (pj.y - pa.y);
            fz += k * (pj.z - pa.z);
            if (VIRIAL) add_virial(ds / d, pj.x - pa.x, pj.y - pa.y, pj.z - pa.z);
        }

        // ---- 3-body
        if (B.n_trios > 0) {
            const int row0 = __ldg(f.off3 + a), n3a = __ldg(f.cnt3 + a);
            const int n_tri = n3a * (n3a - 1) / 2;
            // the centre's neighbours (position, parent atom, species) once per centre in a
            // per-warp table: a triangle then reads two entries instead of gathering two list
            // entries, two positions and two image offsets from global memory
            const bool staged = n3a <= 32;
            const bool tab = staged && leg_table != 0 && NEWTON;
            unsigned char *nb = s_nbr_raw + warp * 32 * NB_REC;
            const unsigned nb_s = smem_addr(nb);
            __syncwarp();       // the previous centre's triangles may still be reading the table
            if (staged && lane < n3a) {
                int aj;
                const Vec3 pj = super_position(f, __ldg(f.idx3 + row0 + lane), aj);
                const unsigned rec = nb_s + NB_REC * (unsigned)lane;
                sts128(rec, make_double2(pj.x, pj.y));
                sts128(rec + 16, make_double2(pj.z, __longlong_as_double(((long long)__ldg(f.spec + aj) << 32) | (unsigned)aj)));
                if (tab) {      // the leg (centre, neighbour), trims applied
                    const double d = dist_rn(pa, pj);
                    const int nkl = __ldg(B_.trio_nk);
                    const double *tl = B.knots3 + __ldg(B_.trio_koff);
                    double v[4] = {0.0, 0.0, 0.0, 0.0}, dv[4] = {0.0, 0.0, 0.0, 0.0};
                    int idx = -1, lo = 0, cnt = 0;
                    if (d >= tl[0] && d <= tl[nkl - 1]) {          // angles.py:502-508
                        idx = eval_leg<false, PS>(tl, nkl, __ldg(B_.trio_scale), B.poly3 + __ldg(B_.trio_poff) / 16 * PS, d,
                                                  B.lead3, B.trail3, v, dv);
                        if (idx >= 0) {
                            lo = max(0, B.lead3 - idx);
                            cnt = max(0, min(4, nkl - 4 - B.trail3 - idx) - lo);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) sts128(rec + 32 + 16 * q, make_double2(v[q], dv[q]));
                    const long long tag = ((long long)(lo | (cnt << 8)) << 32) | (unsigned)idx;
                    sts128(rec + 96, make_double2(d, __longlong_as_double(tag)));
                }
            }
            __syncwarp();
            if (tab) {
                const int nkn = __ldg(B_.trio_nk + 2);
                const double *tn = B.knots3 + __ldg(B_.trio_koff + 2);
                const double *pn = B.poly3 + __ldg(B_.trio_poff + 2) / 16 * PS;
                const double scale_n = __ldg(B_.trio_scale + 2);
                const int dim_m = __ldg(B_.trio_nk + 1) - 4, dim_n = nkn - 4, mn = dim_m * dim_n;
                for (int t = lane; t < n_tri; t += 32) {
                    int qj, qk;
                    unrank_pair(t, qj, qk);
                    const unsigned rj = nb_s + NB_REC * (unsigned)qj, rk = nb_s + NB_REC * (unsigned)qk;
                    const double2 hj = lds128(rj + 96), hk = lds128(rk + 96);
                    const long long tj = __double_as_longlong(hj.y), tk = __double_as_longlong(hk.y);
                    const int il = (int)(tj & 0xffffffffll), im = (int)(tk & 0xffffffffll);
                    const int lo_l = (int)(tj >> 32) & 0xff, cnt_l = (int)(tj >> 40) & 0xff;
                    const int lo_m = (int)(tk >> 32) & 0xff, cnt_m = (int)(tk >> 40) & 0xff;
                    if (il < 0 || im < 0 || cnt_l == 0 || cnt_m == 0) continue;
                    const double2 jxy = lds128(rj), jzt = lds128(rj + 16), kxy = lds128(rk), kzt = lds128(rk + 16);
                    const Vec3 pj = {jxy.x, jxy.y, jzt.x}, pk = {kxy.x, kxy.y, kzt.x};
                    const double djk = dist_rn(pj, pk);
                    if (!(djk >= tn[0] && djk <= tn[nkn - 1])) continue;
                    double vn[4], dvn[4];
                    const int in = eval_leg<false, PS>(tn, nkn, scale_n, pn, djk, B.lead3, B.trail3, vn, dvn);
                    if (in < 0 || in + 4 <= B.lead3 || in >= dim_n - B.trail3) continue;
                    // sum_pqr C[il+p, im+q, in+r] Bl_p Bm_q Bn_r and its three leg-partials over the kept
                    // p, q (trimmed r have zero values)
                    double val = 0.0, gl = 0.0, gm = 0.0, gn = 0.0;
                    const double *base = c_grid + (il * dim_m + im) * dim_n + in;
                    for (int p = lo_l; p < lo_l + cnt_l; ++p) {
                        double u0 = 0.0, u1 = 0.0, u2 = 0.0;
                        for (int q = lo_m; q < lo_m + cnt_m; ++q) {
                            const double *row = base + p * mn + q * dim_n;
                            double t0 = 0.0, t1 = 0.0;
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                const double c = row[r];
                                t0 += c * vn[r];
                                t1 += c * dvn[r];
                            }
                            const double2 m_ = lds128(rk + 32 + 16 * (unsigned)q);     // (B_m, dB_m)
                            u0 += t0 * m_.x;
                            u1 += t0 * m_.y;
                            u2 += t1 * m_.x;
                        }
                        const double2 l_ = lds128(rj + 32 + 16 * (unsigned)p);         // (B_l, dB_l)
                        val += u0 * l_.x;
                        gl += u0 * l_.y;
                        gm += u1 * l_.x;
                        gn += u2 * l_.x;
                    }
                    e_acc += val;
                    const double il_ = fast_rcp(hj.x), im_ = fast_rcp(hk.x), in_ = fast_rcp(djk);
                    const double uij[3] = {(pj.x - pa.x) * il_, (pj.y - pa.y) * il_, (pj.z - pa.z) * il_};
                    const double uik[3] = {(pk.x - pa.x) * im_, (pk.y - pa.y) * im_, (pk.z - pa.z) * im_};
                    const double ujk[3] = {(pk.x - pj.x) * in_, (pk.y - pj.y) * in_, (pk.z - pj.z) * in_};
                    fx += gl * uij[0] + gm * uik[0];
                    fy += gl * uij[1] + gm * uik[1];
                    fz += gl * uij[2] + gm * uik[2];
                    if (VIRIAL) {
                        add_virial(gl * hj.x, uij[0], uij[1], uij[2]);
                        add_virial(gm * hk.x, uik[0], uik[1], uik[2]);
                        add_virial(gn * djk, ujk[0], ujk[1], ujk[2]);
                    }
                    if (want_f) {       // reactions on the parent atoms of j and k
                        const int atom_j = (int)(__double_as_longlong(jzt.y) & 0xffffffffll);
                        const int atom_k = (int)(__double_as_longlong(kzt.y) & 0xffffffffll);
                        double *fj = forces + 3 * (size_t)atom_j, *fk = forces + 3 * (size_t)atom_k;
                        atomicAdd(fj + 0, -gl * uij[0] + gn * ujk[0]);
                        atomicAdd(fj + 1, -gl * uij[1] + gn * ujk[1]);
                        atomicAdd(fj + 2, -gl * uij[2] + gn * ujk[2]);
                        atomicAdd(fk + 0, -gm * uik[0] - gn * ujk[0]);
                        atomicAdd(fk + 1, -gm * uik[1] - gn * ujk[1]);
                        atomicAdd(fk + 2, -gm * uik[2] - gn * ujk[2]);
                    }
                }
            } else
            for (int t = lane; t < n_tri; t += 32) {
                int qj, qk;
                unrank_pair(t, qj, qk);
                Triangle T;
                if (staged) {
                    const unsigned rj = nb_s + NB_REC * (unsigned)qj, rk = nb_s + NB_REC * (unsigned)qk;
                    const double2 jxy = lds128(rj), jzt = lds128(rj + 16), kxy = lds128(rk), kzt = lds128(rk + 16);
                    const long long tj = __double_as_longlong(jzt.y), tk = __double_as_longlong(kzt.y);
                    const Vec3 pj = {jxy.x, jxy.y, jzt.x}, pk = {kxy.x, kxy.y, kzt.x};
                    if (!eval_triangle_at<false, PS>(B, pa, sa, pj, (int)(tj & 0xffffffff), (int)(tj >> 32), pk,
                                                     (int)(tk & 0xffffffff), (int)(tk >> 32), 0, 0, 0, T))
                        continue;
                } else if (!eval_triangle<false, PS>(B, f, pa, sa, __ldg(f.idx3 + row0 + qj), __ldg(f.idx3 + row0 + qk), 0,
                                                     0, 0, T))
                    continue;
                double val, gl, gm, gn;
                contract(c_grid + __ldg(B.trio_goff + T.trio), T, val, gl, gm, gn);
                e_acc += val;
                fx += gl * T.A[0] + gm * T.B[0];
                fy += gl * T.A[1] + gm * T.B[1];
                fz += gl * T.A[2] + gm * T.B[2];
                if (VIRIAL) {       // role 0: A = u_ij, B = u_ik
                    add_virial(gl * T.r[0], T.A[0], T.A[1], T.A[2]);
                    add_virial(gm * T.r[1], T.B[0], T.B[1], T.B[2]);
                    add_virial(gn * T.r[2], T.ujk[0], T.ujk[1], T.ujk[2]);
                }
                if (NEWTON && want_f) {
                    // reactions: F_j = -gl u_ij + gn u_jk, F_k = -gm u_ik - gn u_jk, added to
                    // the parent atoms of j and k; (A, B) hold (u_ij, u_ik), u_jk is rebuilt
                    // from the positions the triangle was evaluated at
                    double *fj = forces + 3 * (size_t)T.atom_j, *fk = forces + 3 * (size_t)T.atom_k;
                    atomicAdd(fj + 0, -gl * T.A[0] + gn * T.ujk[0]);
                    atomicAdd(fj + 1, -gl * T.A[1] + gn * T.ujk[1]);
                    atomicAdd(fj + 2, -gl * T.A[2] + gn * T.ujk[2]);
                    atomicAdd(fk + 0, -gm * T.B[0] - gn * T.ujk[0]);
                    atomicAdd(fk + 1, -gm * T.B[1] - gn * T.ujk[1]);
                    atomicAdd(fk + 2, -gm * T.B[2] - gn * T.ujk[2]);
                }
            }
            if (!NEWTON && want_f) {
                for (int vbase = 0; vbase < n3a; vbase += 32) {
                    const int total = publish_views(B, f, a, vbase, n3a, lane, views);
                    for (int it = lane; it < total; it += 32) {
                        const int v = find_view(views, it);
                        const int ci = views->centre[v], apr = views->a_prime[v];
                        const int mk = __ldg(f.idx3 + __ldg(f.off3 + ci) + (it - views->prefix[v]));
                        if (mk == apr) continue;
                        const bool first = apr < mk;
                        Triangle T;
                        if (!eval_triangle<false, PS>(B, f, real_position(f, ci), __ldg(f.spec + ci), first ? apr : mk,
                                                      first ? mk : apr, first ? 1 : 2, 0, 0, T))
                            continue;
                        double val, gl, gm, gn;
                        contract(c_grid + __ldg(B.trio_goff + T.trio), T, val, gl, gm, gn);
                        fx += gl * T.A[0] + gm * T.B[0] + gn * T.C[0];
                        fy += gl * T.A[1] + gm * T.B[1] + gn * T.C[1];
                        fz += gl * T.A[2] + gm * T.B[2] + gn * T.C[2];
                    }
                    __syncwarp();
                }
            }
        }
        if (want_f) {
            fx = warp_sum(fx);
            fy = warp_sum(fy);
            fz = warp_sum(fz);
            if (lane == 0) {
                if (NEWTON) {       // other centres add their reactions to this atom concurrently
                    atomicAdd(forces + 3 * (size_t)a + 0, fx);
                    atomicAdd(forces + 3 * (size_t)a + 1, fy);
                    atomicAdd(forces + 3 * (size_t)a + 2, fz);
                } else {
                    forces[3 * (size_t)a + 0] = fx;
                    forces[3 * (size_t)a + 1] = fy;
                    forces[3 * (size_t)a + 2] = fz;
                }
            }
        }
    }
    if (want_e) {
        e_acc = warp_sum(e_acc);
        if (lane == 0) e_partials[gw] = e_acc;
    }
    if (VIRIAL) {
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const double s = warp_sum(w[c]);
            if (lane == 0) e_partials[(size_t)n_gw * (1 + c) + gw] = s;
        }
    }
}

// Fixed-order sum of the per-warp energies (block 0) and virial components (blocks 1..6).
__global__ void __launch_bounds__(256) k_energy_sum(const double *__restrict__ partials, int n,
                                                    double *__restrict__ energy) {
    __shared__ double red[256];
    double s = 0.0;
    partials += (size_t)blockIdx.x * n;
    energy += blockIdx.x;
    for (int r = threadIdx.x; r < n; r += blockDim.x) s += partials[r];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int h = 128; h > 0; h >>= 1) {
        if (threadIdx.x < h) red[threadIdx.x] += red[threadIdx.x + h];
        __syncthreads();
    }
    if (threadIdx.x == 0) *energy = red[0];
}

}  // namespace uf3b

using namespace uf3b;

extern "C" int uf3b_energy_forces(uf3b_basis *basis, const uf3b_nlist *nl, double *energy,
                                  double *forces, double *virial, void *stream_) {
    if (!basis || !nl) return fail(UF3B_ERR_INVALID, "null handle");
    if (!basis->has_coeff) return fail(UF3B_ERR_STATE, "coefficients not set");
    if (!energy && !forces && !virial) return UF3B_OK;
    DeviceGuard on_device(basis->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int n = (int)nl->n;
    const bool e_dev = energy && is_device_pointer(energy);
    const bool f_dev = forces && is_device_pointer(forces);
    const bool w_dev = virial && is_device_pointer(virial);
    if (n == 0) {
        if (energy) {
            if (e_dev) UF3B_CUDA(cudaMemsetAsync(energy, 0, sizeof(double), stream));
            else *energy = 0.0;
        }
        if (virial) {
            if (w_dev) UF3B_CUDA(cudaMemsetAsync(virial, 0, 9 * sizeof(double), stream));
            else for (int k = 0; k < 9; ++k) virial[k] = 0.0;
        }
        return UF3B_OK;
    }
    const bool partial = nl->c_count < n;       // this rank owns a range of centres only
    const bool deterministic = getenv("UF3B_DETERMINISTIC_FORCES") != nullptr && !partial;
    const bool newton = (!deterministic && basis->tab.n_trios > 0) || partial;
    const int n_grid = basis->n_bins;
    const size_t grid_bytes = sizeof(double) * (size_t)n_grid;
    const int grid_in_smem = (n_grid > 0 && grid_bytes <= 48 * 1024) ? 1 : 0;
    size_t smem = grid_in_smem ? sizeof(double) * (size_t)((n_grid + 1) & ~1) : 0;
    EvalStage st = {0, 0, 0, 0};
    bool padded = false;
    {   // spline tables next to the coefficient grids (pieces PS_SMEM doubles apart) while the block
        // stays under 56 KB; both or neither, so that one piece stride serves the whole kernel
        const size_t pair_b = sizeof(double) * (size_t)(((basis->n_knots2 + 1) & ~1) + basis->n_poly2 / 16 * PS_SMEM);
        const size_t trio_b = basis->tab.n_trios > 0
            ? sizeof(double) * (size_t)(((basis->n_knots3 + 1) & ~1) + basis->n_poly3 / 16 * PS_SMEM) : 0;
        if (smem + pair_b + trio_b <= 56 * 1024) {
            st.knots2 = basis->n_knots2; st.poly2 = basis->n_poly2;
            if (basis->tab.n_trios > 0) { st.knots3 = basis->n_knots3; st.poly3 = basis->n_poly3; }
            smem += pair_b + trio_b;
            padded = true;
        }
    }
    // leg table path: unary basis whose two centre legs share their knots (featurize_tiled.cu checks the same)
    int leg_table = 0;
    if (newton && basis->tab.n_trios == 1 && basis->tab.ne == 1 && !getenv("UF3B_NO_LEG_TABLE")
        && basis->h_trio_dims[0] == basis->h_trio_dims[1]) {
        leg_table = 1;
        const double *k0 = basis->h_knots3.data() + basis->h_trio_koff[0], *k1 = basis->h_knots3.data() + basis->h_trio_koff[1];
        for (int k = 0; k < basis->h_trio_dims[0] + 4; ++k)
            if (k0[k] != k1[k]) leg_table = 0;
    }
    auto pick = [&](auto pad) {
        constexpr bool P = decltype(pad)::value;
        return virial ? (newton ? k_energy_forces<true, true, P> : k_energy_forces<false, true, P>)
                      : (newton ? k_energy_forces<true, false, P> : k_energy_forces<false, false, P>);
    };
    auto kernel = padded ? pick(std::true_type{}) : pick(std::false_type{});
    UF3B_CUDA(ensure_dynamic_smem((const void *)kernel, smem));
    int per_sm = 1;
    UF3B_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, EV_WARPS * 32, smem));
    if (per_sm < 1) per_sm = 1;
    int grid = sm_count() * per_sm;
    const int need = std::max(1, (nl->c_count + EV_WARPS - 1) / EV_WARPS);
    if (grid > need) grid = need;
    const int n_gw = grid * EV_WARPS;
    UF3B_CUDA(basis->partials.reserve((size_t)n_gw * 7 + 1));
    double *d_f = forces;
    if (forces && !f_dev) {
        UF3B_CUDA(basis->stage.reserve((size_t)3 * n));
        d_f = basis->stage.p;
    }
    double *d_e = energy;
    // sums land in stage_e[0..7) = energy, W_xx, W_yy, W_zz, W_yz, W_xz, W_xy unless the
    // energy alone goes straight to a device address
    if ((energy && !e_dev) || virial) {
        UF3B_CUDA(basis->stage_e.reserve(7));
        d_e = basis->stage_e.p;
    }
    const FrameView view = nl->view();
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (g_timing) {
        UF3B_CUDA(cudaEventCreate(&ev0));
        UF3B_CUDA(cudaEventCreate(&ev1));
        UF3B_CUDA(cudaEventRecord(ev0, stream));
    }
    if (newton && forces) UF3B_CUDA(cudaMemsetAsync(d_f, 0, sizeof(double) * 3 * (size_t)n, stream));
    UF3B_LAUNCH(kernel, grid, EV_WARPS * 32, smem, stream, basis->tab, view, d_f,
                basis->partials.p, energy ? 1 : 0, forces ? 1 : 0, n_grid, grid_in_smem, st, leg_table);
    if (g_timing) UF3B_CUDA(cudaEventRecord(ev1, stream));
    if (energy || virial)
        UF3B_LAUNCH(k_energy_sum, virial ? 7 : 1, 256, 0, stream, basis->partials.p, n_gw, d_e);
    // a deferred list build is verified now, with this call's kernels queued behind it
    if (int rc = nlist_resolve(const_cast<uf3b_nlist *>(nl))) {
        cudaStreamSynchronize(stream);
        return rc;
    }
    bool need_sync = g_timing;
    double h_sums[7] = {0, 0, 0, 0, 0, 0, 0};
    if (virial) {       // 7 doubles: read back, expand the symmetric tensor on the host
        UF3B_CUDA(cudaMemcpyAsync(h_sums, d_e, sizeof(h_sums), cudaMemcpyDeviceToHost, stream));
        if (energy && e_dev)
            UF3B_CUDA(cudaMemcpyAsync(energy, d_e, sizeof(double), cudaMemcpyDeviceToDevice, stream));
        need_sync = true;
    }
    if (forces && !f_dev) {
        UF3B_CUDA(cudaMemcpyAsync(forces, d_f, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, stream));
        need_sync = true;
    }
    if (energy && !e_dev) {
        UF3B_CUDA(cudaMemcpyAsync(energy, d_e, sizeof(double), cudaMemcpyDeviceToHost, stream));
        need_sync = true;
    }
    if (need_sync) UF3B_CUDA(stream_sync(stream));
    if (virial) {
        const double full[9] = {h_sums[1], h_sums[6], h_sums[5], h_sums[6], h_sums[2], h_sums[4],
                                h_sums[5], h_sums[4], h_sums[3]};
        if (w_dev) UF3B_CUDA(cudaMemcpy(virial, full, sizeof(full), cudaMemcpyHostToDevice));
        else for (int k = 0; k < 9; ++k) virial[k] = full[k];
    }
    if (g_timing) {
        float ms = 0.f;
        UF3B_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
        g_last_kernel_ms = ms;
        cudaEventDestroy(ev0);
        cudaEventDestroy(ev1);
    }
    return UF3B_OK;
}
