// featurize_common.cuh — pieces shared by the fit-path kernels (featurize.cu, featurize_tiled.cu).
#pragma once
#include "common.cuh"
#include "geom.cuh"
#include "spline.cuh"

namespace uf3b {

struct __align__(16) PairRec {
    double v[4], dv[4];
    double u[3];
    int col0;              // first feature column hit, or a value no column can match
    int pad;
};

constexpr int CHUNK = 32;

constexpr int ER_SPLIT = 32;      // intermediate rows of the two-stage energy-row sum

// 2-body rows of atom `a` (bspline.py:810-895) added into acc[4 * col + (e, fx, fy, fz)]:
// lanes evaluate 32 pairs at a time into `prec`, then every lane gathers the records that
// touch ITS feature column.
// `chunk0`, `chunk_step`: the passes (of 32 pairs) this caller takes — all of them by default;
// the cooperative kernel deals even and odd passes to two warps with separate accumulators.
__device__ __forceinline__ void two_body_rows(const BasisTab &B, const FrameView &f, int a, int sa, const Vec3 &pa,
                                              double *acc, PairRec *prec, int lane, int chunk0 = 0,
                                              int chunk_step = 1) {
    const int r0 = __ldg(f.off2 + a), r1 = r0 + __ldg(f.cnt2 + a);
    for (int base = r0 + chunk0 * CHUNK; base < r1; base += chunk_step * CHUNK) {
        const int e = base + lane;
        PairRec rec;
        rec.col0 = -(1 << 20);
        if (e < r1) {
            int aj;
            const Vec3 pj = super_position(f, __ldg(f.idx2 + e), aj);
            const double d = dist_rn(pa, pj);
            const int pr = pair_index(B.ne, sa, __ldg(f.spec + aj));
            const int idx = eval_leg(B.knots2 + __ldg(B.pair_koff + pr), __ldg(B.pair_nk + pr), __ldg(B.pair_scale + pr),
                                     B.poly2 + __ldg(B.pair_poff + pr), d, B.lead2, B.trail2,
                                     rec.v, rec.dv);
            if (idx >= 0) {
                rec.col0 = __ldg(B.pair_col + pr) + idx;
                const double inv = 1.0 / d;
                rec.u[0] = (pj.x - pa.x) * inv;
                rec.u[1] = (pj.y - pa.y) * inv;
                rec.u[2] = (pj.z - pa.z) * inv;
            }
        }
        prec[lane] = rec;
        __syncwarp();
        const int count = min(CHUNK, r1 - base);
        for (int sj = 0; sj < B.ne; ++sj) {
            const int pr = pair_index(B.ne, sa, sj);
            const int c0 = __ldg(B.pair_col + pr), nb = __ldg(B.pair_nk + pr) - 4;
            for (int cb = 0; cb < nb; cb += 32) {
                const int col = c0 + cb + lane;
                double se = 0.0, sx = 0.0, sy = 0.0, sz = 0.0;
                for (int t = 0; t < count; ++t) {
                    const unsigned r = (unsigned)(col - prec[t].col0);
                    if (r < 4u) {
                        se += prec[t].v[r];
                        const double dv = prec[t].dv[r];
                        sx += dv * prec[t].u[0];
                        sy += dv * prec[t].u[1];
                        sz += dv * prec[t].u[2];
                    }
                }
                if (cb + lane < nb) {
                    // every bond is seen from both ends (distances.py:118-120):
                    // x[a] = 2 * sum_j B'(r_aj) (x_j - x_a) / r_aj
                    double *dst = acc + 4 * (size_t)col;
                    dst[0] += se;
                    dst[1] += 2.0 * sx;
                    dst[2] += 2.0 * sy;
                    dst[3] += 2.0 * sz;
                }
            }
        }
        __syncwarp();
    }
}


// partials [n_rows][F] (+ ER_SPLIT scratch rows behind them) -> energy row d_xe
int launch_energy_row(double *partials, int n_rows, int F, double *d_xe, cudaStream_t stream);

// Copies to host buffers (if any), synchronisation and kernel timing shared by the launch paths.
int finish_featurize(uf3b_basis *basis, double *x_energy, double *x_forces, int64_t ld, double *d_xe,
                     double *d_xf, int F, int n, bool e_dev, bool f_dev, cudaStream_t stream,
                     cudaEvent_t ev0, cudaEvent_t ev1);

// featurize_tiled.cu: takes the frame when the basis fits the register-tiled kernel, else returns 1
int featurize_tiled(uf3b_basis *basis, const uf3b_nlist *nl, double *x_energy, double *x_forces, int64_t ld,
                    cudaStream_t stream, bool deferred);

// featurize_multi.cu: several species / symmetry-1 trios / long rows; returns 1 when it does not apply
int featurize_multi(uf3b_basis *basis, const uf3b_nlist *nl, double *x_energy, double *x_forces, int64_t ld,
                    cudaStream_t stream);

}  // namespace uf3b
