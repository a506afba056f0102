/*
 * oracle/uf3_oracle.c — CPU restatement of the UF3 featurization / evaluator hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The product (uf3_b200/) never does.
 *
 * It restates, in plain scalar C (IEEE float64, no FMA contraction: build with
 * -ffp-contract=off), what the pure-Python reference computes with dense numpy/scipy
 * arrays, following the reference's own formulation (explicit ghost supercell, ghost
 * centred triangles in the force path) so that the CUDA path's reformulation
 * (real centres only, contributions routed to parent atoms) is tested against it
 * rather than against itself.  Pinned against the running reference and its golden
 * vectors by tests/test_oracle_golden.py (fixtures: tests/golden/, generator:
 * oracle/make_golden.py).
 *
 * Reference lines restated (all relative to /root/reference/uf3):
 *   supercell            data/geometry.py:141-149 (image_rank*N + atom; offsets from host)
 *   pair list, energy    representation/distances.py:19-75   (strict r_min < d < r_max)
 *   pair list, forces    representation/distances.py:78-143, :331-364
 *   2-body features      representation/bspline.py:810-849, :852-895
 *   3-body neighbours    representation/angles.py:289-346    (r_min < d <= r_max)
 *   triplets             representation/angles.py:424-514    (j<k, ghost rule, Z order, leg filter)
 *   interval lookup      representation/bspline.py:950-974   (searchsorted 'left' - 4)
 *   leg evaluation       representation/angles.py:517-632    (trimmed indices stay 0)
 *   scatter              representation/angles.py:104-139, :235-286
 *   compression          representation/bspline.py:664-690   (as a bin -> column map)
 *   evaluator            forcefield/calculator.py:183-343
 *
 * O(N) through a cell list over the explicit supercell instead of cdist's O(M^2)
 * matrix; distances use the same expression as scipy's cdist:
 * sqrt(((dx*dx) + dy*dy) + dz*dz).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t n_elements;
    const int32_t *numbers;      /* [n_elements] ascending atomic numbers                */
    int32_t n_pairs;
    const int32_t *pair_nk;      /* [n_pairs] knots per pair                             */
    const int32_t *pair_koff;    /* [n_pairs] offset into knots2                         */
    const int32_t *pair_foff;    /* [n_pairs] first feature column of the pair           */
    const double *knots2;
    const double *pair_rmin;     /* [n_pairs] r_min_map (clamped at 0 when used)         */
    const double *pair_rmax;
    int32_t lead2, trail2, lead3, trail3;
    int32_t n_trios;
    const int32_t *trio_nk;      /* [n_trios*3]                                          */
    const int32_t *trio_koff;    /* [n_trios*3] offsets into knots3                      */
    const int32_t *trio_foff;    /* [n_trios] first feature column                       */
    const int32_t *trio_goff;    /* [n_trios] offset of the trio's L*M*N block in bin_*  */
    const double *knots3;
    const int32_t *bin_col;      /* full-grid bin -> compressed column (-1: dropped)     */
    const double *bin_w;         /* weight applied when folding the bin into the column  */
    double r3min, r3max;         /* angles.py:312-325                                    */
    double r_cut;
    int32_t n_feats;
} orc_basis;

typedef struct {
    int64_t n_real, n_sup;
    double *pos;                 /* [n_sup*3] */
    int32_t *spec;               /* [n_sup] element index */
    /* cell list over the whole supercell */
    double lo[3], edge;
    int32_t dim[3];
    int64_t *cell_start;         /* [ncell+1] */
    int64_t *cell_items;         /* [n_sup]   */
} orc_super;

static int pair_index(int ne, int a, int b)
{
    if (a > b) { int t = a; a = b; b = t; }
    return a * ne - a * (a - 1) / 2 + (b - a);
}

static double dist(const double *p, const double *q)
{
    double dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
    return sqrt(dx * dx + dy * dy + dz * dz);
}

/* ---------------------------------------------------------------- supercell */
static void super_free(orc_super *s)
{
    free(s->pos); free(s->spec); free(s->cell_start); free(s->cell_items);
    memset(s, 0, sizeof *s);
}

static int cell_of(const orc_super *s, const double *p, int c[3])
{
    for (int d = 0; d < 3; ++d) {
        int v = (int)floor((p[d] - s->lo[d]) / s->edge);
        if (v < 0) v = 0;
        if (v >= s->dim[d]) v = s->dim[d] - 1;
        c[d] = v;
    }
    return (c[0] * s->dim[1] + c[1]) * s->dim[2] + c[2];
}

/* geometry.py:141-149 — ghost = position + image offset; index = image*N + atom. */
static int super_build(orc_super *s, const orc_basis *b, int64_t n, const double *pos,
                       const int32_t *spec, int64_t n_img, const double *offsets)
{
    memset(s, 0, sizeof *s);
    s->n_real = n;
    s->n_sup = n * n_img;
    s->pos = (double *)malloc(sizeof(double) * 3 * (size_t)(s->n_sup ? s->n_sup : 1));
    s->spec = (int32_t *)malloc(sizeof(int32_t) * (size_t)(s->n_sup ? s->n_sup : 1));
    if (!s->pos || !s->spec) return -1;
    for (int64_t g = 0; g < n_img; ++g)
        for (int64_t a = 0; a < n; ++a) {
            int64_t m = g * n + a;
            for (int d = 0; d < 3; ++d) s->pos[3 * m + d] = pos[3 * a + d] + offsets[3 * g + d];
            s->spec[m] = spec[a];
        }
    double hi[3] = {0, 0, 0};
    for (int d = 0; d < 3; ++d) { s->lo[d] = 0; hi[d] = 0; }
    for (int64_t m = 0; m < s->n_sup; ++m)
        for (int d = 0; d < 3; ++d) {
            double v = s->pos[3 * m + d];
            if (m == 0 || v < s->lo[d]) s->lo[d] = v;
            if (m == 0 || v > hi[d]) hi[d] = v;
        }
    s->edge = b->r_cut > b->r3max ? b->r_cut : b->r3max;
    if (!(s->edge > 0)) s->edge = 1.0;
    s->edge *= 1.0 + 1e-9;
    int64_t ncell = 1;
    for (int d = 0; d < 3; ++d) {
        double span = hi[d] - s->lo[d];
        int64_t k = (int64_t)floor(span / s->edge) + 1;
        if (k > 512) { k = 512; }
        s->dim[d] = (int32_t)k;
        ncell *= k;
    }
    /* if the grid was capped, widen the cells of that axis by growing edge uniformly */
    for (int d = 0; d < 3; ++d) {
        double span = hi[d] - s->lo[d];
        if (span / s->edge >= s->dim[d]) s->edge = span / s->dim[d] * (1.0 + 1e-9);
    }
    s->cell_start = (int64_t *)calloc((size_t)ncell + 1, sizeof(int64_t));
    s->cell_items = (int64_t *)malloc(sizeof(int64_t) * (size_t)(s->n_sup ? s->n_sup : 1));
    if (!s->cell_start || !s->cell_items) return -1;
    int c[3];
    for (int64_t m = 0; m < s->n_sup; ++m) s->cell_start[cell_of(s, s->pos + 3 * m, c) + 1]++;
    for (int64_t k = 0; k < ncell; ++k) s->cell_start[k + 1] += s->cell_start[k];
    int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (size_t)ncell);
    if (!fill) return -1;
    memcpy(fill, s->cell_start, sizeof(int64_t) * (size_t)ncell);
    for (int64_t m = 0; m < s->n_sup; ++m)   /* ascending m => items sorted within a cell */
        s->cell_items[fill[cell_of(s, s->pos + 3 * m, c)]++] = m;
    free(fill);
    return 0;
}

typedef struct { int64_t *idx; double *d; int64_t n, cap; } orc_row;

static int row_push(orc_row *r, int64_t j, double d)
{
    if (r->n == r->cap) {
        int64_t cap = r->cap ? 2 * r->cap : 64;
        int64_t *ni = (int64_t *)realloc(r->idx, sizeof(int64_t) * (size_t)cap);
        double *nd = (double *)realloc(r->d, sizeof(double) * (size_t)cap);
        if (!ni || !nd) return -1;
        r->idx = ni; r->d = nd; r->cap = cap;
    }
    r->idx[r->n] = j; r->d[r->n] = d; r->n++;
    return 0;
}

static void row_sort(orc_row *r)       /* insertion sort by supercell index */
{
    for (int64_t a = 1; a < r->n; ++a) {
        int64_t ji = r->idx[a]; double dd = r->d[a]; int64_t b = a - 1;
        while (b >= 0 && r->idx[b] > ji) { r->idx[b + 1] = r->idx[b]; r->d[b + 1] = r->d[b]; --b; }
        r->idx[b + 1] = ji; r->d[b + 1] = dd;
    }
}

/* mode 2: pair bounds per interaction, strict both sides (distances.py:60-66,129-134).
 * mode 3: r3min < d <= r3max regardless of species (angles.py:340,344).           */
static int neighbours(const orc_super *s, const orc_basis *b, int64_t i, int mode, orc_row *out)
{
    out->n = 0;
    const double *pi = s->pos + 3 * i;
    int c[3];
    cell_of(s, pi, c);
    for (int x = c[0] - 1; x <= c[0] + 1; ++x) {
        if (x < 0 || x >= s->dim[0]) continue;
        for (int y = c[1] - 1; y <= c[1] + 1; ++y) {
            if (y < 0 || y >= s->dim[1]) continue;
            for (int z = c[2] - 1; z <= c[2] + 1; ++z) {
                if (z < 0 || z >= s->dim[2]) continue;
                int64_t cell = ((int64_t)x * s->dim[1] + y) * s->dim[2] + z;
                for (int64_t q = s->cell_start[cell]; q < s->cell_start[cell + 1]; ++q) {
                    int64_t j = s->cell_items[q];
                    double d = dist(pi, s->pos + 3 * j);
                    int keep;
                    if (mode == 2) {
                        int p = pair_index(b->n_elements, s->spec[i], s->spec[j]);
                        double lo = b->pair_rmin[p] > 0 ? b->pair_rmin[p] : 0;
                        keep = (d > lo) && (d < b->pair_rmax[p]);
                    } else {
                        keep = (d > b->r3min) && (d <= b->r3max);
                    }
                    if (keep && row_push(out, j, d)) return -1;
                }
            }
        }
    }
    row_sort(out);
    return 0;
}

/* -------------------------------------------------------------- B-splines */
/* numpy.searchsorted(knots, r, side='left'): first index with knots[idx] >= r. */
static int searchsorted_left(const double *t, int n, double r)
{
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) / 2; if (t[mid] < r) lo = mid + 1; else hi = mid; }
    return lo;
}

/* The four cubic B-splines that are non-zero on (t[i], t[i+1]] and their first
 * derivatives (Cox-de Boor; the same quantities scipy's basis_element(...)(x, nu)
 * returns for basis indices i-3..i).  `i` is the knot interval, 3 <= i <= n-5.   */
static void bspline4(const double *t, int i, double x, double v[4], double dv[4])
{
    double N[4][4];
    double left[4], right[4];
    N[0][0] = 1.0;
    for (int j = 1; j <= 3; ++j) {
        left[j] = x - t[i + 1 - j];
        right[j] = t[i + j] - x;
        double saved = 0.0;
        for (int r = 0; r < j; ++r) {
            double den = right[r + 1] + left[j - r];
            N[j][r] = den;                      /* lower triangle keeps knot differences */
            double temp = den != 0.0 ? N[r][j - 1] / den : 0.0;
            N[r][j] = saved + right[r + 1] * temp;
            saved = left[j - r] * temp;
        }
        N[j][j] = saved;
    }
    for (int r = 0; r < 4; ++r) v[r] = N[r][3];
    /* first derivative: 3 * (B_{r-1,2}/(t_{r+3}-t_r) - B_{r,2}/(t_{r+4}-t_{r+1}))   */
    for (int r = 0; r < 4; ++r) {
        double a = 0.0, c = 0.0;
        if (r >= 1) { double den = N[3][r - 1]; a = den != 0.0 ? N[r - 1][2] / den : 0.0; }
        if (r <= 2) { double den = N[3][r];     c = den != 0.0 ? N[r][2] / den : 0.0; }
        dv[r] = 3.0 * (a - c);
    }
}

/* interval + values with trims applied: returns first basis index or -1 when r is
 * outside (t[0], t[-1]].  Entries whose basis index is outside
 * [n_lead, n_basis - n_trail) are zeroed (angles.py:554-565; bspline.py:840).    */
static int eval_leg(const double *t, int nk, double r, int n_lead, int n_trail,
                    double v[4], double dv[4])
{
    int nb = nk - 4;
    int idx = searchsorted_left(t, nk, r) - 4;
    if (idx < 0 || idx > nb - 4) return -1;
    bspline4(t, idx + 3, r, v, dv);
    for (int q = 0; q < 4; ++q) {
        int bi = idx + q;
        if (bi < n_lead || bi >= nb - n_trail) { v[q] = 0.0; dv[q] = 0.0; }
    }
    return idx;
}

/* ---------------------------------------------------------------- exports */
void orc_free(void *p) { free(p); }

/* Parity hook: CSR neighbour lists for the real centres.  mode 2 or 3 as above. */
int orc_neighbor_lists(const orc_basis *b, int64_t n, const double *pos, const int32_t *spec,
                       int64_t n_img, const double *offsets, int mode,
                       int64_t **offsets_out, int64_t **idx_out)
{
    orc_super s;
    if (super_build(&s, b, n, pos, spec, n_img, offsets)) return -1;
    int64_t *off = (int64_t *)calloc((size_t)n + 1, sizeof(int64_t));
    int64_t cap = 1024, cnt = 0;
    int64_t *idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap);
    orc_row row = {0};
    for (int64_t i = 0; i < n; ++i) {
        if (neighbours(&s, b, i, mode, &row)) return -1;
        while (cnt + row.n > cap) { cap *= 2; idx = (int64_t *)realloc(idx, sizeof(int64_t) * (size_t)cap); }
        memcpy(idx + cnt, row.idx, sizeof(int64_t) * (size_t)row.n);
        cnt += row.n;
        off[i + 1] = cnt;
    }
    free(row.idx); free(row.d);
    super_free(&s);
    *offsets_out = off; *idx_out = idx;
    return 0;
}

/* One triangle (i; j, k) already in reference leg order. */
typedef struct {
    int t;                 /* trio interaction */
    int il, im, in;        /* first basis index per leg */
    double v[3][4], dv[3][4];
    double rl, rm, rn;
} orc_tri;

/* angles.py:480-512: order (j,k) by atomic number (stable), pick the interaction,
 * apply the inclusive leg filter, evaluate the legs.  Returns 0 if the triangle
 * contributes nothing.                                                          */
static int make_triangle(const orc_super *s, const orc_basis *b, int64_t i, int64_t *j, int64_t *k,
                         double dij, double dik, orc_tri *tri)
{
    if (s->spec[*j] > s->spec[*k]) { int64_t t = *j; *j = *k; *k = t; double d = dij; dij = dik; dik = d; }
    int ne = b->n_elements;
    int t = s->spec[i] * b->n_pairs + pair_index(ne, s->spec[*j], s->spec[*k]);
    const int32_t *nk = b->trio_nk + 3 * t;
    const double *tl = b->knots3 + b->trio_koff[3 * t + 0];
    const double *tm = b->knots3 + b->trio_koff[3 * t + 1];
    const double *tn = b->knots3 + b->trio_koff[3 * t + 2];
    double djk = dist(s->pos + 3 * *j, s->pos + 3 * *k);
    if (!(dij >= tl[0] && dij <= tl[nk[0] - 1])) return 0;
    if (!(dik >= tm[0] && dik <= tm[nk[1] - 1])) return 0;
    if (!(djk >= tn[0] && djk <= tn[nk[2] - 1])) return 0;
    tri->t = t; tri->rl = dij; tri->rm = dik; tri->rn = djk;
    tri->il = eval_leg(tl, nk[0], dij, b->lead3, b->trail3, tri->v[0], tri->dv[0]);
    tri->im = eval_leg(tm, nk[1], dik, b->lead3, b->trail3, tri->v[1], tri->dv[1]);
    tri->in = eval_leg(tn, nk[2], djk, b->lead3, b->trail3, tri->v[2], tri->dv[2]);
    if (tri->il < 0 || tri->im < 0 || tri->in < 0) return 0;   /* r exactly on knots[0] */
    return 1;
}

/*
 * Feature rows for one configuration (process.py:293-367 without the y column):
 *   x_energy [n_feats]          = [n_el counts, 2-body, 3-body]
 *   x_forces [3*n, n_feats]     row c*n + a   (fx_0.. fx_{n-1}, fy_0.., fz_0..)
 * Either pointer may be NULL to skip that part.
 */
int orc_featurize(const orc_basis *b, int64_t n, const double *pos, const int32_t *spec,
                  int64_t n_img, const double *offsets, double *x_energy, double *x_forces)
{
    orc_super s;
    if (super_build(&s, b, n, pos, spec, n_img, offsets)) return -1;
    const int F = b->n_feats;
    const int ne = b->n_elements;
    orc_row row = {0};
    double v[4], dv[4];
    if (x_energy) memset(x_energy, 0, sizeof(double) * (size_t)F);
    if (x_forces) memset(x_forces, 0, sizeof(double) * (size_t)F * 3 * (size_t)n);

    if (x_energy) {
        /* composition.py:96-111 */
        for (int64_t a = 0; a < n; ++a) x_energy[spec[a]] += 1.0;
        /* distances.py:19-75 + bspline.py:810-849 */
        for (int64_t i = 0; i < n; ++i) {
            if (neighbours(&s, b, i, 2, &row)) return -1;
            for (int64_t q = 0; q < row.n; ++q) {
                int p = pair_index(ne, s.spec[i], s.spec[row.idx[q]]);
                int idx = eval_leg(b->knots2 + b->pair_koff[p], b->pair_nk[p], row.d[q],
                                   b->lead2, b->trail2, v, dv);
                if (idx < 0) continue;
                for (int r = 0; r < 4; ++r) x_energy[b->pair_foff[p] + idx + r] += v[r];
            }
        }
    }
    if (x_forces) {
        /* distances.py:78-143,331-364 + bspline.py:852-895: every ordered pair (i, j) of
         * the supercell with i or j real; x[m,c,b] = -sum_p B'_b(r_p)(d(m,j)-d(m,i))(x_j-x_i)_c/r_p */
        for (int64_t i = 0; i < s.n_sup; ++i) {
            if (neighbours(&s, b, i, 2, &row)) return -1;
            for (int64_t q = 0; q < row.n; ++q) {
                int64_t j = row.idx[q];
                if (i >= n && j >= n) continue;
                int p = pair_index(ne, s.spec[i], s.spec[j]);
                int idx = eval_leg(b->knots2 + b->pair_koff[p], b->pair_nk[p], row.d[q],
                                   b->lead2, b->trail2, v, dv);
                if (idx < 0) continue;
                for (int c = 0; c < 3; ++c) {
                    double delta = (s.pos[3 * j + c] - s.pos[3 * i + c]) / row.d[q];
                    for (int r = 0; r < 4; ++r) {
                        int col = b->pair_foff[p] + idx + r;
                        if (j < n) x_forces[((size_t)c * n + j) * F + col] -= dv[r] * delta;
                        if (i < n) x_forces[((size_t)c * n + i) * F + col] += dv[r] * delta;
                    }
                }
            }
        }
    }
    if (b->n_trios > 0) {
        /* angles.py:17-78 (energy: real centres) and :142-232 (forces: every centre) */
        int64_t n_centres = x_forces ? s.n_sup : n;
        orc_tri tri;
        for (int64_t i = 0; i < n_centres; ++i) {
            if (neighbours(&s, b, i, 3, &row)) return -1;
            for (int64_t qj = 0; qj < row.n; ++qj) {
                if (i >= n && row.idx[qj] >= n) continue;        /* ghost centre: j must be real */
                for (int64_t qk = 0; qk < row.n; ++qk) {
                    if (!(row.idx[qj] < row.idx[qk])) continue;
                    int64_t j = row.idx[qj], k = row.idx[qk];
                    if (!make_triangle(&s, b, i, &j, &k, row.d[qj], row.d[qk], &tri)) continue;
                    const int32_t *nk = b->trio_nk + 3 * tri.t;
                    int M = nk[1] - 4, N = nk[2] - 4;
                    const int32_t *bc = b->bin_col + b->trio_goff[tri.t];
                    const double *bw = b->bin_w + b->trio_goff[tri.t];
                    int foff = b->trio_foff[tri.t];
                    /* direction cosines, distances.py:354-363 */
                    double uij[3], uik[3], ujk[3];
                    for (int c = 0; c < 3; ++c) {
                        uij[c] = (s.pos[3 * j + c] - s.pos[3 * i + c]) / tri.rl;
                        uik[c] = (s.pos[3 * k + c] - s.pos[3 * i + c]) / tri.rm;
                        ujk[c] = (s.pos[3 * k + c] - s.pos[3 * j + c]) / tri.rn;
                    }
                    for (int p = 0; p < 4; ++p)
                        for (int q = 0; q < 4; ++q)
                            for (int r = 0; r < 4; ++r) {
                                int bin = ((tri.il + p) * M + (tri.im + q)) * N + (tri.in + r);
                                int col = bc[bin];
                                if (col < 0) continue;
                                double w = bw[bin];
                                if (x_energy && i < n)
                                    x_energy[foff + col] += w * tri.v[0][p] * tri.v[1][q] * tri.v[2][r];
                                if (!x_forces) continue;
                                double gl = tri.dv[0][p] * tri.v[1][q] * tri.v[2][r];
                                double gm = tri.v[0][p] * tri.dv[1][q] * tri.v[2][r];
                                double gn = tri.v[0][p] * tri.v[1][q] * tri.dv[2][r];
                                for (int c = 0; c < 3; ++c) {
                                    /* angles.py:282-285 with kronecker (m==j)-(m==i) etc., then
                                     * force_grids -= grids (angles.py:229-231)            */
                                    if (i < n) x_forces[((size_t)c * n + i) * F + foff + col]
                                                   -= w * (-gl * uij[c] - gm * uik[c]);
                                    if (j < n) x_forces[((size_t)c * n + j) * F + foff + col]
                                                   -= w * (gl * uij[c] - gn * ujk[c]);
                                    if (k < n) x_forces[((size_t)c * n + k) * F + foff + col]
                                                   -= w * (gm * uik[c] + gn * ujk[c]);
                                }
                            }
                }
            }
        }
    }
    free(row.idx); free(row.d);
    super_free(&s);
    return 0;
}

/*
 * Evaluator (calculator.py:156-343).  coeff: flat model.coefficients [n_feats];
 * c_grid: per-trio decompressed coefficient grids, concatenated like bin_col
 * (bspline.py:693-719).  The 2-body spline is NOT trimmed at evaluation
 * (calculator.py:207,286 call NDSpline on every coefficient).
 */
int orc_energy_forces(const orc_basis *b, const double *coeff, const double *c_grid,
                      int64_t n, const double *pos, const int32_t *spec,
                      int64_t n_img, const double *offsets, double *energy, double *forces)
{
    orc_super s;
    if (super_build(&s, b, n, pos, spec, n_img, offsets)) return -1;
    const int ne = b->n_elements;
    orc_row row = {0};
    double v[4], dv[4];
    if (energy) {
        double e = 0.0;
        for (int64_t a = 0; a < n; ++a) e += coeff[spec[a]];            /* :183-189 */
        for (int64_t i = 0; i < n; ++i) {                               /* :191-211 */
            if (neighbours(&s, b, i, 2, &row)) return -1;
            for (int64_t q = 0; q < row.n; ++q) {
                int p = pair_index(ne, s.spec[i], s.spec[row.idx[q]]);
                int idx = eval_leg(b->knots2 + b->pair_koff[p], b->pair_nk[p], row.d[q], 0, 0, v, dv);
                if (idx < 0) continue;
                for (int r = 0; r < 4; ++r) e += coeff[b->pair_foff[p] + idx + r] * v[r];
            }
        }
        *energy = e;
    }
    if (forces) {
        memset(forces, 0, sizeof(double) * 3 * (size_t)n);
        for (int64_t i = 0; i < s.n_sup; ++i) {                         /* :267-291 */
            if (neighbours(&s, b, i, 2, &row)) return -1;
            for (int64_t q = 0; q < row.n; ++q) {
                int64_t j = row.idx[q];
                if (i >= n && j >= n) continue;
                int p = pair_index(ne, s.spec[i], s.spec[j]);
                int idx = eval_leg(b->knots2 + b->pair_koff[p], b->pair_nk[p], row.d[q], 0, 0, v, dv);
                if (idx < 0) continue;
                double ds = 0.0;
                for (int r = 0; r < 4; ++r) ds += coeff[b->pair_foff[p] + idx + r] * dv[r];
                for (int c = 0; c < 3; ++c) {
                    double delta = (s.pos[3 * j + c] - s.pos[3 * i + c]) / row.d[q];
                    if (j < n) forces[3 * j + c] -= ds * delta;
                    if (i < n) forces[3 * i + c] += ds * delta;
                }
            }
        }
    }
    if (b->n_trios > 0) {
        int64_t n_centres = forces ? s.n_sup : n;                       /* :213-244, :293-343 */
        orc_tri tri;
        /* the evaluator's NDSpline applies no trims: the zeros live in c_grid */
        orc_basis nb = *b; nb.lead3 = 0; nb.trail3 = 0;
        for (int64_t i = 0; i < n_centres; ++i) {
            if (neighbours(&s, b, i, 3, &row)) return -1;
            for (int64_t qj = 0; qj < row.n; ++qj) {
                if (i >= n && row.idx[qj] >= n) continue;
                for (int64_t qk = 0; qk < row.n; ++qk) {
                    if (!(row.idx[qj] < row.idx[qk])) continue;
                    int64_t j = row.idx[qj], k = row.idx[qk];
                    if (!make_triangle(&s, &nb, i, &j, &k, row.d[qj], row.d[qk], &tri)) continue;
                    const int32_t *nk = b->trio_nk + 3 * tri.t;
                    int M = nk[1] - 4, N = nk[2] - 4;
                    const double *cg = c_grid + b->trio_goff[tri.t];
                    double val = 0, gl = 0, gm = 0, gn = 0;
                    for (int p = 0; p < 4; ++p)
                        for (int q = 0; q < 4; ++q)
                            for (int r = 0; r < 4; ++r) {
                                double cc = cg[((tri.il + p) * M + (tri.im + q)) * N + (tri.in + r)];
                                val += cc * tri.v[0][p] * tri.v[1][q] * tri.v[2][r];
                                gl += cc * tri.dv[0][p] * tri.v[1][q] * tri.v[2][r];
                                gm += cc * tri.v[0][p] * tri.dv[1][q] * tri.v[2][r];
                                gn += cc * tri.v[0][p] * tri.v[1][q] * tri.dv[2][r];
                            }
                    if (energy && i < n) *energy += val;
                    if (!forces) continue;
                    for (int c = 0; c < 3; ++c) {
                        double uij = (s.pos[3 * j + c] - s.pos[3 * i + c]) / tri.rl;
                        double uik = (s.pos[3 * k + c] - s.pos[3 * i + c]) / tri.rm;
                        double ujk = (s.pos[3 * k + c] - s.pos[3 * j + c]) / tri.rn;
                        if (i < n) forces[3 * i + c] -= -gl * uij - gm * uik;
                        if (j < n) forces[3 * j + c] -= gl * uij - gn * ujk;
                        if (k < n) forces[3 * k + c] -= gm * uik + gn * ujk;
                    }
                }
            }
        }
    }
    free(row.idx); free(row.d);
    super_free(&s);
    return 0;
}

/* Standalone spline probe for unit tests: values/derivatives of the four non-zero
 * basis functions at r (no trims); returns the first basis index or -1.          */
int orc_eval_basis(const double *knots, int32_t nk, double r, double v[4], double dv[4])
{
    return eval_leg(knots, nk, r, 0, 0, v, dv);
}
