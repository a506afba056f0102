// neighbors.cu — Kernel A: neighbour lists over a cell-linked grid of the ghost supercell.
//
// Replaces geometry.get_supercell + distances.get_distance_matrix + the boolean masks of
// the reference (data/geometry.py:14-51; representation/distances.py:19-75,146-169;
// representation/angles.py:289-346), which materialise a dense (atoms x supercell)
// distance matrix.  Here every periodic image that can lie within the search radius of
// the real atoms is binned into cells of edge >= r_search (counting sort, 32-byte slots,
// z-runs of cells contiguous), one thread block walks one cell, one warp one real centre,
// and ragged per-centre hit counts are compacted with warp ballots.  Two CSR lists are
// produced per centre, both sorted by supercell index (image_rank * n_atoms + atom):
//   list 2:  max(r_min,0) < d < r_max  of the pair's own bounds    (distances.py:60-66)
//   list 3:  r3min < d <= r3max                                      (angles.py:340)
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>

#include "common.cuh"
#include "geom.cuh"

namespace uf3b {

struct GridParams {
    double ox, oy, oz;     // lower corner of the binned region
    double hx, hy, hz;     // upper corner (ghosts outside are dropped)
    double edge;
    int nx, ny, nz;
};

__device__ __forceinline__ void atomic_min_f64(double *addr, double v) {
    unsigned long long *p = reinterpret_cast<unsigned long long *>(addr);
    unsigned long long old = *p, assumed;
    do {
        assumed = old;
        if (__longlong_as_double((long long)assumed) <= v) break;
        old = atomicCAS(p, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
}
__device__ __forceinline__ void atomic_max_f64(double *addr, double v) {
    unsigned long long *p = reinterpret_cast<unsigned long long *>(addr);
    unsigned long long old = *p, assumed;
    do {
        assumed = old;
        if (__longlong_as_double((long long)assumed) >= v) break;
        old = atomicCAS(p, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
}

// species lookup for every atom + bounding box of the CENTRES [c_first, c_first + c_count)
// (all real atoms unless the frame is split over ranks: the grid only has to hold what lies
// within the search radius of this rank's centres) in bbox[0..5] and of ALL atoms in bbox[8..13]
// (the binning kernels skip whole periodic images with it).  Every block leaves its twelve bounds
// in `partial`; the last block to finish (ticket counter, reset for the next launch) folds them —
// one min / max atomic per bound and block on twelve shared addresses serialised in the L2 (the
// kernel took 14-28 us for 10 000 atoms that way).
constexpr int PREP_THREADS = 256;
__global__ void __launch_bounds__(PREP_THREADS) k_prepare(int n, int c_first, int c_count, const int *__restrict__ z,
                                                          const int *__restrict__ z_to_spec,
                                                          const double *__restrict__ pos,
                                                          int *__restrict__ spec, double *__restrict__ bbox,
                                                          int *__restrict__ err, double *__restrict__ partial,
                                                          unsigned *__restrict__ ticket) {
    double v[12] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY,
                    INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int bad = 0;
    for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < n; a += gridDim.x * blockDim.x) {
        const int zz = z[a];
        const int s = (zz >= 0 && zz < 128) ? z_to_spec[zz] : -1;
        spec[a] = s < 0 ? 0 : s;
        if (s < 0) bad = 1;
        const bool centre = (unsigned)(a - c_first) < (unsigned)c_count;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double x = pos[3 * a + c];
            v[6 + c] = fmin(v[6 + c], x);
            v[9 + c] = fmax(v[9 + c], x);
            if (centre) { v[c] = fmin(v[c], x); v[3 + c] = fmax(v[3 + c], x); }
        }
    }
    __shared__ double red[12][PREP_THREADS / 32];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto fold = [&](int c, double x) {          // warp reduction of bound c
        const bool is_min = (c % 6) < 3;
        for (int s = 16; s > 0; s >>= 1) {
            const double o = __shfl_xor_sync(FULL, x, s);
            x = is_min ? fmin(x, o) : fmax(x, o);
        }
        return x;
    };
#pragma unroll
    for (int c = 0; c < 12; ++c) {
        const double x = fold(c, v[c]);
        if (lane == 0) red[c][warp] = x;
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(err, 1);
    if (warp == 0) {
#pragma unroll
        for (int c = 0; c < 12; ++c) {
            const bool is_min = (c % 6) < 3;
            const double x = fold(c, lane < PREP_THREADS / 32 ? red[c][lane] : (is_min ? INFINITY : -INFINITY));
            if (lane == 0) partial[12 * blockIdx.x + c] = x;
        }
        if (lane == 0) {
            __threadfence();
            last = atomicAdd(ticket, 1u) == gridDim.x - 1;
        }
    }
    __syncthreads();
    if (!last || warp != 0) return;
    __threadfence();
    // all twelve bounds of a block's partial row are loaded together (one bound at a time the loop was a
    // chain of twelve dependent L2 round trips: 10 of the kernel's 15 us)
    double t[12];
#pragma unroll
    for (int c = 0; c < 12; ++c) t[c] = (c % 6) < 3 ? INFINITY : -INFINITY;
    for (int b = lane; b < (int)gridDim.x; b += 32) {
        double o[12];
#pragma unroll
        for (int c = 0; c < 12; ++c) o[c] = __ldcg(partial + 12 * b + c);
#pragma unroll
        for (int c = 0; c < 12; ++c) t[c] = (c % 6) < 3 ? fmin(t[c], o[c]) : fmax(t[c], o[c]);
    }
#pragma unroll
    for (int c = 0; c < 12; ++c) {
        const double x = fold(c, t[c]);
        if (lane == 0) bbox[c < 6 ? c : c + 2] = x;         // centres at [0..5], all atoms at [8..13]
    }
    if (lane == 0) *ticket = 0;
}

__device__ __forceinline__ int cell_coord(double p, double o, double edge, int n) {
    int c = (int)floor((p - o) / edge);
    return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

// A periodic image none of whose atoms can lie in the binned region: the box of ALL atoms
// (bbox[8..13], k_prepare) shifted by the image offset misses the grid.  With the centres of one
// rank (a slab of the frame) almost every image of a large cell goes this way without its
// positions being read.
__device__ __forceinline__ bool image_outside(const double *__restrict__ bbox, const double *__restrict__ off,
                                              const GridParams &G) {
    return bbox[8] + off[0] > G.hx || bbox[11] + off[0] < G.ox || bbox[9] + off[1] > G.hy || bbox[12] + off[1] < G.oy
           || bbox[10] + off[2] > G.hz || bbox[13] + off[2] < G.oz;
}

// One thread per supercell atom m = g*n + a (blockIdx.y = image g): drop ghosts outside the padded
// bounding box of the centres, count the rest per cell.  Blocks of an image that cannot reach the
// grid leave at once (cell_of is only read for the images that stay: k_bin_fill repeats the test).
__global__ void k_bin_count(int n, int n_img, const double *__restrict__ pos,
                            const double *__restrict__ img_off, GridParams G, const double *__restrict__ bbox,
                            int *__restrict__ cell_of, int *__restrict__ cell_cnt) {
  for (int g = blockIdx.y; g < n_img; g += gridDim.y) {
    if (g != 0 && image_outside(bbox, img_off + 3 * g, G)) continue;
    const double ox = img_off[3 * g + 0], oy = img_off[3 * g + 1], oz = img_off[3 * g + 2];
    for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < n; a += gridDim.x * blockDim.x) {
        const double x = __dadd_rn(pos[3 * a + 0], ox);
        const double y = __dadd_rn(pos[3 * a + 1], oy);
        const double z = __dadd_rn(pos[3 * a + 2], oz);
        int cell = -1;
        if (g == 0 || (x >= G.ox && x <= G.hx && y >= G.oy && y <= G.hy && z >= G.oz && z <= G.hz)) {
            const int cx = cell_coord(x, G.ox, G.edge, G.nx);
            const int cy = cell_coord(y, G.oy, G.edge, G.ny);
            const int cz = cell_coord(z, G.oz, G.edge, G.nz);
            cell = (cx * G.ny + cy) * G.nz + cz;
            atomicAdd(cell_cnt + cell, 1);
        }
        cell_of[(long long)g * n + a] = cell;
    }
  }
}

__global__ void k_bin_fill(int n, int n_img, const double *__restrict__ pos,
                           const double *__restrict__ img_off, GridParams G, const double *__restrict__ bbox,
                           const int *__restrict__ spec,
                           const int *__restrict__ cell_of, const int *__restrict__ cell_start,
                           int *__restrict__ cursor, Slot *__restrict__ slots) {
  for (int g = blockIdx.y; g < n_img; g += gridDim.y) {
    if (g != 0 && image_outside(bbox, img_off + 3 * g, G)) continue;
    const double ox = img_off[3 * g + 0], oy = img_off[3 * g + 1], oz = img_off[3 * g + 2];
    for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < n; a += gridDim.x * blockDim.x) {
        const long long m = (long long)g * n + a;
        const int cell = cell_of[m];
        if (cell < 0) continue;
        Slot s;
        s.x = __dadd_rn(pos[3 * a + 0], ox);
        s.y = __dadd_rn(pos[3 * a + 1], oy);
        s.z = __dadd_rn(pos[3 * a + 2], oz);
        s.m = (int)m;
        s.spec = spec[a];
        slots[cell_start[cell] + atomicAdd(cursor + cell, 1)] = s;
    }
  }
}

// Exclusive scan of up to two int arrays (block b scans array b); out[n] = total.
constexpr int SCAN_ITEMS = 8;
__global__ void __launch_bounds__(1024) k_scan(const int *in0, int *out0, const int *in1,
                                               int *out1, int n, long long *totals) {
    const int *in = blockIdx.x ? in1 : in0;
    int *out = blockIdx.x ? out1 : out0;
    __shared__ long long warp_tot[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long carry = 0;
    for (long long base = 0; base < n; base += 1024 * SCAN_ITEMS) {
        const long long i0 = base + (long long)threadIdx.x * SCAN_ITEMS;
        int vals[SCAN_ITEMS];
        long long local = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            vals[k] = (i0 + k < n) ? in[i0 + k] : 0;
            local += vals[k];
        }
        long long inc = local;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const long long up = __shfl_up_sync(FULL, inc, s);
            if (lane >= s) inc += up;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            long long w = warp_tot[lane];
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const long long up = __shfl_up_sync(FULL, w, s);
                if (lane >= s) w += up;
            }
            warp_tot[lane] = w;
        }
        __syncthreads();
        long long excl = carry + inc - local + (warp ? warp_tot[warp - 1] : 0);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            if (i0 + k < n) out[i0 + k] = (int)excl;
            excl += vals[k];
        }
        carry += warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[n] = (int)carry;
        totals[blockIdx.x] = carry;
    }
}

// Three-kernel scan for long arrays (blockIdx.y selects the array): tile-local exclusive
// scans + tile totals, a one-block scan of the totals, then the tile offsets are added.
constexpr int SCAN_TILE = 1024 * SCAN_ITEMS;
__global__ void __launch_bounds__(1024) k_scan_local(const int *in0, int *out0, const int *in1, int *out1,
                                                     int n, long long *tile_sums, int n_tiles) {
    const int *in = blockIdx.y ? in1 : in0;
    int *out = blockIdx.y ? out1 : out0;
    __shared__ long long warp_tot[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long i0 = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    int vals[SCAN_ITEMS];
    long long local = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        vals[k] = (i0 + k < n) ? in[i0 + k] : 0;
        local += vals[k];
    }
    long long inc = local;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const long long up = __shfl_up_sync(FULL, inc, s);
        if (lane >= s) inc += up;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        long long w = warp_tot[lane];
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const long long up = __shfl_up_sync(FULL, w, s);
            if (lane >= s) w += up;
        }
        warp_tot[lane] = w;
    }
    __syncthreads();
    long long excl = inc - local + (warp ? warp_tot[warp - 1] : 0);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (i0 + k < n) out[i0 + k] = (int)excl;
        excl += vals[k];
    }
    if (threadIdx.x == 0) tile_sums[(size_t)blockIdx.y * n_tiles + blockIdx.x] = warp_tot[31];
}

__global__ void __launch_bounds__(1024) k_scan_tiles(long long *tile_sums, int n_tiles, int *out0, int *out1,
                                                     int n, long long *totals) {
    long long *sums = tile_sums + (size_t)blockIdx.x * n_tiles;
    int *out = blockIdx.x ? out1 : out0;
    __shared__ long long warp_tot[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long v = (int)threadIdx.x < n_tiles ? sums[threadIdx.x] : 0;
    long long inc = v;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const long long up = __shfl_up_sync(FULL, inc, s);
        if (lane >= s) inc += up;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        long long w = warp_tot[lane];
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const long long up = __shfl_up_sync(FULL, w, s);
            if (lane >= s) w += up;
        }
        warp_tot[lane] = w;
    }
    __syncthreads();
    if ((int)threadIdx.x < n_tiles) sums[threadIdx.x] = inc - v + (warp ? warp_tot[warp - 1] : 0);
    if (threadIdx.x == 0) {
        out[n] = (int)warp_tot[31];
        totals[blockIdx.x] = warp_tot[31];
    }
}

__global__ void __launch_bounds__(1024) k_scan_add(int *out0, int *out1, int n, const long long *tile_sums,
                                                   int n_tiles) {
    int *out = blockIdx.y ? out1 : out0;
    const long long off = tile_sums[(size_t)blockIdx.y * n_tiles + blockIdx.x];
    const long long i0 = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (i0 + k < n) out[i0 + k] += (int)off;
}

// Row of `n` distinct ints: out[rank(v)] = v.
__device__ __forceinline__ void warp_rank_sort(const int *src, int *dst, int n, int lane) {
    for (int e = lane; e < n; e += 32) {
        const int v = src[e];
        int rank = 0;
        for (int k = 0; k < n; ++k) rank += src[k] < v;
        dst[rank] = v;
    }
}

constexpr int NL_WARPS = 4;
constexpr int NL_TILE = 1024;     // slots staged per block (32 KB); larger tiles fall back to global reads

constexpr int NL_BUF2 = 160, NL_BUF3 = 96;     // hits buffered per centre before the row is placed
constexpr int NL_REGIONS = 128;                // independent claim regions per list
constexpr int NL_PAIRS = 36;                   // pair bounds staged in shared memory (up to 8 elements)

// Block = one cell, warp = one real centre of that cell.  The slots of the 3x3x3 block of
// cells around it are 9 contiguous z-runs in the binned array; one thread stages them into
// shared memory with 9 TMA bulk copies tracked by an mbarrier, and every warp then sweeps
// the staged tile 32 candidates at a time.  One pass: hits are ballot-compacted into a
// per-warp shared buffer, the row's place in the index array is claimed with one atomicAdd
// per list, and the row is written sorted by supercell index.  Rows are contiguous but
// appear in claim order: the lists are (start, count) per centre.  If the arrays are too
// small the claims still count, `status[2]` is raised and the host grows them and reruns.
// Claims go to NL_REGIONS independent regions of the index arrays (region = block % NL_REGIONS,
// one counter each): with a single counter per list the 10^5 same-address atomics of a frame
// serialised in the L2 and the warps spent 41 % of the kernel waiting for their claim.
//   claims: [r] entries claimed in region r of list 2, [NL_REGIONS + r] of list 3
//   status: [2] overflow, [3] longest list-3 row   ([0], [1], [4], [5] are filled by
//           k_post_status: totals of both lists and the fullest region of each)
__global__ void __launch_bounds__(NL_WARPS * 32)
k_neighbors(const BasisTab B, const GridParams G, const Slot *__restrict__ slots,
            const int *__restrict__ cell_start, int c_first, int c_count, int *__restrict__ off2,
            int *__restrict__ cnt2, int *__restrict__ off3, int *__restrict__ cnt3,
            int *__restrict__ idx2, int *__restrict__ idx3, int *__restrict__ scratch2,
            int *__restrict__ scratch3, int cap2, int cap3, int *__restrict__ status,
            int *__restrict__ claims) {
    __shared__ __align__(128) Slot tile[NL_TILE];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ int run_lo[9], run_n[9], n_cand;
    __shared__ int hit2[NL_WARPS][NL_BUF2], hit3[NL_WARPS][NL_BUF3];
    __shared__ double s_lo[NL_PAIRS], s_hi[NL_PAIRS];     // strict pair bounds (global beyond NL_PAIRS)
    const int cell = blockIdx.x;
    const int s0 = cell_start[cell], s1 = cell_start[cell + 1];
    if (s0 == s1) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // any real atom in this cell?  (ghost-only cells of the padding do no work)
    int real_here = 0;
    for (int s = s0 + (int)threadIdx.x; s < s1; s += NL_WARPS * 32) real_here |= (unsigned)(slots[s].m - c_first) < (unsigned)c_count;
    if (!__syncthreads_or(real_here)) return;

    const bool bounds_staged = B.n_pairs <= NL_PAIRS;
    if (bounds_staged)
        for (int p = threadIdx.x; p < B.n_pairs; p += NL_WARPS * 32) { s_lo[p] = B.pair_lo[p]; s_hi[p] = B.pair_hi[p]; }
    const double *pair_lo = bounds_staged ? s_lo : B.pair_lo, *pair_hi = bounds_staged ? s_hi : B.pair_hi;
    const int cz = cell % G.nz, cy = (cell / G.nz) % G.ny, cx = cell / (G.nz * G.ny);
    if (threadIdx.x < 32) {       // warp 0: lane k < 9 owns z-run k = (dx, dy) of the 3 x 3 x 3 block
        const int k = threadIdx.x;
        int lo = 0, cnt = 0;
        if (k < 9) {
            const int x = cx + k / 3 - 1, y = cy + k % 3 - 1;
            if (x >= 0 && x < G.nx && y >= 0 && y < G.ny) {
                const int z0 = cz > 0 ? cz - 1 : 0, z1 = cz + 1 < G.nz ? cz + 1 : G.nz - 1;
                const int row = (x * G.ny + y) * G.nz;
                lo = cell_start[row + z0];
                cnt = cell_start[row + z1 + 1] - lo;
            }
            run_lo[k] = lo;
            run_n[k] = cnt;
        }
        int before = cnt;                 // inclusive prefix of the run lengths over lanes 0..8
        for (int sft = 1; sft < 16; sft <<= 1) {
            const int up = __shfl_up_sync(FULL, before, sft);
            if (k >= sft) before += up;
        }
        const int total = __shfl_sync(FULL, before, 8);
        before -= cnt;
        const unsigned b = smem_addr(&bar);
        if (k == 0) {
            n_cand = total;
            if (total <= NL_TILE) {
                mbar_init(b, 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                mbar_expect_tx(b, (unsigned)total * (unsigned)sizeof(Slot));
            }
        }
        __syncwarp();
        if (total <= NL_TILE && k < 9 && cnt > 0)
            bulk_copy_g2s(smem_addr(tile) + (unsigned)before * (unsigned)sizeof(Slot), slots + lo,
                          (unsigned)cnt * (unsigned)sizeof(Slot), b);
    }
    __syncthreads();
    const int total = n_cand;
    const bool staged = total <= NL_TILE;
    if (staged) mbar_wait(smem_addr(&bar), 0);

    const unsigned lt = (1u << lane) - 1u;
    const bool has3 = B.n_trios > 0;
    const int region = cell % NL_REGIONS;
    const double r2_search = B.r_search * B.r_search * (1.0 + 1e-12);     // no cutoff exceeds r_search
    for (int s = s0 + warp; s < s1; s += NL_WARPS) {
        const Slot c = slots[s];
        if ((unsigned)(c.m - c_first) >= (unsigned)c_count) continue;   // ghosts (and real atoms of
                                                                       // other ranks) are never centres
        const Vec3 pc = {c.x, c.y, c.z};
        // pass = 0: count and buffer; pass = 1 (only for rows longer than the buffers): write
        // the hits straight to the global scratch rows
        int n2 = 0, n3 = 0, base2 = 0, base3 = 0;
        bool fits = true, placed = true;
        for (int pass = 0; pass < 2; ++pass) {
            n2 = n3 = 0;
            auto visit = [&](const Slot *cand, int count) {
                for (int q0 = 0; q0 < count; q0 += 32) {
                    const int q = q0 + lane;
                    bool k2 = false, k3 = false;
                    int m = 0;
                    if (q < count) {
                        const Slot t = cand[q];
                        // squared distance first (the same rounded operations dist_rn takes the
                        // root of): candidates beyond the search radius need no square root
                        const double dx = __dsub_rn(pc.x, t.x), dy = __dsub_rn(pc.y, t.y), dz = __dsub_rn(pc.z, t.z);
                        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                        if (d2 <= r2_search) {
                            const double d = __dsqrt_rn(d2);
                            const int p = pair_index(B.ne, c.spec, t.spec);
                            k2 = d > pair_lo[p] && d < pair_hi[p];
                            k3 = has3 && d > B.r3min && d <= B.r3max;
                        }
                        m = t.m;
                    }
                    const unsigned b2 = __ballot_sync(FULL, k2), b3 = __ballot_sync(FULL, k3);
                    const int p2 = n2 + __popc(b2 & lt), p3 = n3 + __popc(b3 & lt);
                    if (pass == 0) {
                        if (k2 && p2 < NL_BUF2) hit2[warp][p2] = m;
                        if (k3 && p3 < NL_BUF3) hit3[warp][p3] = m;
                    } else if (placed) {
                        if (k2) scratch2[base2 + p2] = m;
                        if (k3) scratch3[base3 + p3] = m;
                    }
                    n2 += __popc(b2);
                    n3 += __popc(b3);
                }
            };
            if (staged) {
                visit(tile, total);
            } else {
                for (int r = 0; r < 9; ++r) visit(slots + run_lo[r], run_n[r]);
            }
            if (pass == 1) break;
            // claim the rows
            if (lane == 0) {
                base2 = atomicAdd(claims + region, n2);
                base3 = atomicAdd(claims + NL_REGIONS + region, n3);
                atomicMax(status + 3, n3);
            }
            base2 = __shfl_sync(FULL, base2, 0);
            base3 = __shfl_sync(FULL, base3, 0);
            placed = base2 + n2 <= cap2 && base3 + n3 <= cap3;     // cap: entries per region
            base2 += region * cap2;
            base3 += region * cap3;
            if (!placed && lane == 0) status[2] = 1;
            fits = n2 <= NL_BUF2 && n3 <= NL_BUF3;
            if (fits || !placed) break;
        }
        __syncwarp();
        if (lane == 0) {
            off2[c.m] = base2; cnt2[c.m] = placed ? n2 : 0;
            off3[c.m] = base3; cnt3[c.m] = placed ? n3 : 0;
        }
        if (placed) {
            if (fits) {
                warp_rank_sort(hit2[warp], idx2 + base2, n2, lane);
                warp_rank_sort(hit3[warp], idx3 + base3, n3, lane);
            } else {
                warp_rank_sort(scratch2 + base2, idx2 + base2, n2, lane);
                warp_rank_sort(scratch3 + base3, idx3 + base3, n3, lane);
            }
        }
        __syncwarp();
    }
}

}  // namespace uf3b

using namespace uf3b;

// Exclusive scan of one or two int arrays of length n on `stream` (out[n] = total).
static int scan_arrays(uf3b_nlist *nl, int n_arrays, const int *in0, int *out0, const int *in1, int *out1,
                       int n, long long *totals, cudaStream_t stream) {
    const int n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (n_tiles <= 1 || n_tiles > 1024) {
        UF3B_LAUNCH(k_scan, n_arrays, 1024, 0, stream, in0, out0, in1, out1, n, totals);
        return UF3B_OK;
    }
    UF3B_CUDA(nl->tile_sums.reserve(2 * (size_t)n_tiles));
    UF3B_LAUNCH(k_scan_local, dim3(n_tiles, n_arrays), 1024, 0, stream, in0, out0, in1, out1, n,
                nl->tile_sums.p, n_tiles);
    UF3B_LAUNCH(k_scan_tiles, n_arrays, 1024, 0, stream, nl->tile_sums.p, n_tiles, out0, out1, n, totals);
    UF3B_LAUNCH(k_scan_add, dim3(n_tiles, n_arrays), 1024, 0, stream, out0, out1, n, nl->tile_sums.p, n_tiles);
    return UF3B_OK;
}

FrameView uf3b_nlist::view() const {
    FrameView f;
    f.n = (int)n;
    f.n_magic = n > 1 ? (unsigned)((1ull << 32) / (unsigned long long)n) : 0xffffffffu;
    f.n_img = n_img;
    f.c_first = c_first;
    f.c_count = c_count;
    f.pos = pos.p;
    f.spec = spec.p;
    f.img_off = img_off.p;
    f.img_inv = img_inv.p;
    f.off2 = off2.p; f.cnt2 = cnt2.p; f.idx2 = idx2.p;
    f.off3 = off3.p; f.cnt3 = cnt3.p; f.idx3 = idx3.p;
    return f;
}

// n 8-byte words device -> device-mapped host memory (see uf3b_nlist::h_mapped)
__global__ void k_post_small(const double *__restrict__ src, volatile double *dst, int n) {
    if ((int)threadIdx.x < n) dst[threadIdx.x] = src[threadIdx.x];
    __threadfence_system();
}

// totals and fullest region of both lists -> status[0], [1], [4], [5]; status -> mapped host memory
// status[6]: the centres left the box (+ skin) the reused grid was sized for; status[7]: element /
// position error flags of k_prepare (only checked here when the grid was reused)
struct BoxCheck { int active; double lo[3], hi[3], skin; };

// status[8]: 1 when a consumer queued behind a DEFERRED build must not use it — overflow, stale grid,
// bad input, or a 3-body row longer than `max3_hint` (the longest row of the previous build, which the
// feature kernels size their shared memory by when they cannot wait for this build's own value).
__global__ void k_post_status(const int *__restrict__ claims, int *__restrict__ status, volatile int *dst,
                              const double *__restrict__ misc, const BoxCheck chk, int max3_hint) {
    const int lane = threadIdx.x;
    if (chk.active && lane == 0) {
        int stale = 0, bad = ((const int *)(misc + 6))[0] ? 1 : 0;
        for (int c = 0; c < 3; ++c) {
            const double lo = misc[c], hi = misc[3 + c];
            if (!(lo >= chk.lo[c] - chk.skin) || !(hi <= chk.hi[c] + chk.skin)) stale = 1;
            if (!isfinite(lo) || !isfinite(hi)) bad |= 2;
        }
        status[6] = stale;
        status[7] = bad;
    }
    __syncwarp();
    long long t2 = 0, t3 = 0;
    int m2 = 0, m3 = 0;
    for (int r = lane; r < NL_REGIONS; r += 32) {
        const int a = claims[r], b = claims[NL_REGIONS + r];
        t2 += a; t3 += b;
        m2 = max(m2, a); m3 = max(m3, b);
    }
    for (int s = 16; s > 0; s >>= 1) {
        t2 += __shfl_xor_sync(FULL, t2, s); t3 += __shfl_xor_sync(FULL, t3, s);
        m2 = max(m2, __shfl_xor_sync(FULL, m2, s)); m3 = max(m3, __shfl_xor_sync(FULL, m3, s));
    }
    if (lane == 0) {
        status[0] = t2 > INT32_MAX ? -1 : (int)t2;
        status[1] = t3 > INT32_MAX ? -1 : (int)t3;
        status[4] = m2;
        status[5] = m3;
    }
    if (lane == 0)
        status[8] = (status[2] != 0 || status[6] != 0 || status[7] != 0 || status[3] > max3_hint) ? 1 : 0;
    __syncwarp();
    if (lane < 9) dst[lane] = status[lane];
    __threadfence_system();
}

int uf3b::nlist_resolve(uf3b_nlist *nl) {
    if (!nl || !nl->pending) return UF3B_OK;
    nl->pending = false;
    UF3B_CUDA(cudaEventSynchronize(nl->status_ev));
    int h_status[8];
    memcpy(h_status, nl->h_mapped + 8, sizeof h_status);
    if (h_status[7] & 1) return fail(UF3B_ERR_ELEMENT, "configuration holds an element outside the basis");
    if (h_status[7] & 2) return fail(UF3B_ERR_INVALID, "non-finite position");
    if (h_status[0] < 0 || h_status[1] < 0) return fail(UF3B_ERR_CAPACITY, "neighbour list exceeds int32 offsets");
    if (h_status[6]) {              // the centres left the cached box: the next build takes a fresh grid
        nl->grid_valid = false;
        return fail(UF3B_RETRY, "deferred list build: the atoms left the cached cell grid");
    }
    if (h_status[2]) {              // index arrays too small: wait for whoever reads them, then regrow
        UF3B_CUDA(cudaStreamSynchronize(nl->pending_stream));
        UF3B_CUDA(nl->idx2.reserve((size_t)(h_status[4] + 1) * NL_REGIONS));
        UF3B_CUDA(nl->idx3.reserve((size_t)(h_status[5] + 1) * NL_REGIONS));
        return fail(UF3B_RETRY, "deferred list build: index arrays regrown");
    }
    nl->total2 = h_status[0];
    nl->total3 = h_status[1];
    nl->max3 = h_status[3];
    if (nl->max3 > nl->max3_hint) {     // a consumer that sized itself by the hint skipped this frame
        nl->max3_hint = nl->max3;
        if (nl->hint_used) return fail(UF3B_RETRY, "deferred list build: a 3-body row outgrew the previous frame's longest");
    }
    nl->max3_hint = nl->max3;
    return UF3B_OK;
}

extern "C" {

int uf3b_neighbors_build(uf3b_basis *basis, int64_t n_atoms, const double *positions,
                         const int32_t *atomic_numbers, int32_t n_images,
                         const double *image_offsets, const int32_t *image_abc,
                         uf3b_nlist **inout, void *stream_) {
    return uf3b_neighbors_build_range(basis, n_atoms, positions, atomic_numbers, n_images, image_offsets,
                                      image_abc, 0, n_atoms, inout, stream_);
}

int uf3b_neighbors_build_range(uf3b_basis *basis, int64_t n_atoms, const double *positions,
                               const int32_t *atomic_numbers, int32_t n_images,
                               const double *image_offsets, const int32_t *image_abc,
                               int64_t first_centre, int64_t n_centres,
                               uf3b_nlist **inout, void *stream_) {
    if (!basis || !inout) return fail(UF3B_ERR_INVALID, "null argument");
    if (first_centre < 0 || n_centres < 0 || first_centre + n_centres > n_atoms)
        return fail(UF3B_ERR_INVALID, "centre range outside [0, n_atoms)");
    if (n_atoms < 0 || n_images < 1) return fail(UF3B_ERR_INVALID, "bad n_atoms / n_images");
    if (n_atoms > 0 && (!positions || !atomic_numbers)) return fail(UF3B_ERR_INVALID, "null positions");
    if (!image_offsets || !image_abc) return fail(UF3B_ERR_INVALID, "null image table");
    const long long n_sup = (long long)n_atoms * n_images;
    if (n_sup >= (1LL << 31) - 64) return fail(UF3B_ERR_CAPACITY, "supercell exceeds int32 indices");
    DeviceGuard on_device(basis->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool created = (*inout == nullptr);
    uf3b_nlist *nl = created ? new uf3b_nlist() : *inout;
    struct Guard {   // frees a list created by a failing call
        uf3b_nlist *p; bool armed;
        ~Guard() { if (armed) delete p; }
    } guard{nl, created};

    if (nl->pending) {          // an unverified deferred build is superseded by this one ...
        const bool consumed = nl->consumed_pending;
        nl->consumed_pending = false;
        const int rc = nlist_resolve(nl);
        if (rc < 0 && rc != UF3B_ERR_CAPACITY) { guard.armed = false; return rc; }
        // ... unless feature rows were produced from it (uf3b_featurize with device outputs does not wait
        // for the build): an invalid build then means invalid rows, and the caller has to hear about it
        if (rc == UF3B_RETRY && consumed) {
            guard.armed = false;
            return fail(UF3B_RETRY, "the previous frame's deferred list build was invalid: its feature rows must "
                                    "not be used; repeat that frame, then this build");
        }
    }
    const int n = (int)n_atoms;
    nl->n = n_atoms;
    nl->n_img = n_images;
    nl->total2 = nl->total3 = 0;
    nl->max3 = 0;
    nl->c_first = (int)first_centre;
    nl->c_count = (int)n_centres;

    // periodic image table: pair every image with the one of negated coordinates.  An unchanged table (MD
    // loops, streams of frames of one cell) is already on the device: no inversion map, no two copies
    const bool same_images = nl->img_uploaded == n_images && !is_device_pointer(image_offsets)
                             && nl->img_host.size() == 3 * (size_t)n_images
                             && memcmp(nl->img_host.data(), image_offsets, sizeof(double) * 3 * n_images) == 0
                             && nl->abc_host.size() == 3 * (size_t)n_images
                             && memcmp(nl->abc_host.data(), image_abc, sizeof(int32_t) * 3 * n_images) == 0;
    std::vector<int> inv(same_images ? 0 : n_images, 0);
    if (!same_images) {
        std::map<std::tuple<int, int, int>, int> rank;
        for (int g = 0; g < n_images; ++g)
            rank[std::make_tuple(image_abc[3 * g], image_abc[3 * g + 1], image_abc[3 * g + 2])] = g;
        if (image_abc[0] || image_abc[1] || image_abc[2])
            return fail(UF3B_ERR_INVALID, "image 0 must be the home cell");
        for (int g = 0; g < n_images; ++g) {
            auto it = rank.find(std::make_tuple(-image_abc[3 * g], -image_abc[3 * g + 1], -image_abc[3 * g + 2]));
            if (it == rank.end()) return fail(UF3B_ERR_INVALID, "image table is not inversion symmetric");
            inv[g] = it->second;
        }
    }
    UF3B_CUDA(nl->img_off.reserve(3 * (size_t)n_images));
    UF3B_CUDA(nl->img_inv.reserve(n_images));
    UF3B_CUDA(nl->off2.reserve((size_t)n + 1));
    UF3B_CUDA(nl->off3.reserve((size_t)n + 1));
    UF3B_CUDA(nl->totals.reserve(8 + NL_REGIONS));      // [0] scan total, [1..4] status (8 ints), [8..] claims
    if (!same_images) {
        UF3B_CUDA(cudaMemcpyAsync(nl->img_off.p, image_offsets, sizeof(double) * 3 * n_images, cudaMemcpyDefault, stream));
        UF3B_CUDA(cudaMemcpyAsync(nl->img_inv.p, inv.data(), sizeof(int) * n_images, cudaMemcpyHostToDevice, stream));
        nl->img_uploaded = 0;
        if (!is_device_pointer(image_offsets)) {
            // pageable sources are staged by the runtime before the call returns: `inv` may go out of scope
            nl->img_host.assign(image_offsets, image_offsets + 3 * (size_t)n_images);
            nl->abc_host.assign(image_abc, image_abc + 3 * (size_t)n_images);
            nl->img_uploaded = n_images;
        }
    }
    if (n == 0) {
        UF3B_CUDA(stream_sync(stream));
        guard.armed = false;
        *inout = nl;
        return UF3B_OK;
    }
    UF3B_CUDA(nl->pos.reserve(3 * (size_t)n));
    UF3B_CUDA(nl->z.reserve(n));
    UF3B_CUDA(nl->spec.reserve(n));
    constexpr int PREP_BLOCKS_MAX = 256;
    UF3B_CUDA(nl->misc.reserve(16 + 12 * PREP_BLOCKS_MAX + 2));
    UF3B_CUDA(cudaMemcpyAsync(nl->pos.p, positions, sizeof(double) * 3 * n, cudaMemcpyDefault, stream));
    UF3B_CUDA(cudaMemcpyAsync(nl->z.p, atomic_numbers, sizeof(int) * n, cudaMemcpyDefault, stream));
    int *d_err = (int *)(nl->misc.p + 6);
    // k_prepare's last block writes all twelve bounds; only the error words have to be cleared
    UF3B_CUDA(cudaMemsetAsync(nl->misc.p + 6, 0, 2 * sizeof(double), stream));
    const int prep_blocks = std::min((n + PREP_THREADS - 1) / PREP_THREADS, PREP_BLOCKS_MAX);
    if (!nl->ticket_zeroed) {       // the kernel leaves the counter at zero for the next launch
        UF3B_CUDA(cudaMemsetAsync(nl->misc.p + 16 + 12 * PREP_BLOCKS_MAX, 0, 2 * sizeof(double), stream));
        nl->ticket_zeroed = true;
    }
    UF3B_LAUNCH(k_prepare, prep_blocks, PREP_THREADS, 0, stream, n, nl->c_first, nl->c_count, nl->z.p, basis->tab.z_to_spec, nl->pos.p,
                nl->spec.p, nl->misc.p, d_err, nl->misc.p + 16, (unsigned *)(nl->misc.p + 16 + 12 * PREP_BLOCKS_MAX));
    if (!nl->h_mapped) UF3B_CUDA(cudaHostAlloc((void **)&nl->h_mapped, 16 * sizeof(double), cudaHostAllocMapped));
    // reuse the previous build's grid (no bounding-box read-back) when the problem is the same
    // and — checked on the device, read back with the list totals — the centres moved less
    // than the skin it was padded with
    const double grid_skin = 1.0;
    static const bool no_reuse = getenv("UF3B_NO_GRID_REUSE") != nullptr;
    const bool reuse = !no_reuse && nl->grid_valid && nl->grid_n == n_atoms && nl->c_count > 0
                       && nl->grid_rsearch == basis->tab.r_search
                       && nl->grid_cfirst == nl->c_first && nl->grid_ccount == nl->c_count
                       && nl->grid_imgoff.size() == 3 * (size_t)n_images && !is_device_pointer(image_offsets)
                       && memcmp(nl->grid_imgoff.data(), image_offsets, sizeof(double) * 3 * n_images) == 0;
    double h_misc[7] = {0, 0, 0, 0, 0, 0, 0};
    int h_err = 0;
    if (!reuse) {
        UF3B_LAUNCH(k_post_small, 1, 32, 0, stream, nl->misc.p, nl->h_mapped, 7);
        UF3B_CUDA(stream_sync(stream));
        memcpy(h_misc, nl->h_mapped, sizeof h_misc);
        memcpy(&h_err, &h_misc[6], sizeof h_err);
    }
    nl->grid_valid = false;
    if (h_err) return fail(UF3B_ERR_ELEMENT, "configuration holds an element outside the basis");
    if (nl->c_count == 0) {         // a rank without centres: empty rows, nothing to bin
        UF3B_CUDA(nl->cnt2.reserve((size_t)n + 1));
        UF3B_CUDA(nl->cnt3.reserve((size_t)n + 1));
        UF3B_CUDA(cudaMemsetAsync(nl->cnt2.p, 0, sizeof(int) * n, stream));
        UF3B_CUDA(cudaMemsetAsync(nl->cnt3.p, 0, sizeof(int) * n, stream));
        UF3B_CUDA(cudaMemsetAsync(nl->off2.p, 0, sizeof(int) * n, stream));
        UF3B_CUDA(cudaMemsetAsync(nl->off3.p, 0, sizeof(int) * n, stream));
        if (nl->idx2.cap == 0) UF3B_CUDA(nl->idx2.reserve(64));
        if (nl->idx3.cap == 0) UF3B_CUDA(nl->idx3.reserve(64));
        guard.armed = false;
        *inout = nl;
        return UF3B_OK;
    }
    GridParams G;
    long long n_cell;
    if (reuse) {
        G.ox = nl->grid_par[0]; G.oy = nl->grid_par[1]; G.oz = nl->grid_par[2];
        G.hx = nl->grid_par[3]; G.hy = nl->grid_par[4]; G.hz = nl->grid_par[5];
        G.edge = nl->grid_par[6];
        G.nx = nl->grid_dim[0]; G.ny = nl->grid_dim[1]; G.nz = nl->grid_dim[2];
        n_cell = (long long)G.nx * G.ny * G.nz;
    } else {
        for (int c = 0; c < 6; ++c)
            if (!std::isfinite(h_misc[c])) return fail(UF3B_ERR_INVALID, "non-finite position");
        // grid over the bounding box of the centres padded by the search radius and the skin
        const double pad = basis->tab.r_search * (1.0 + 1e-9);
        G.edge = pad;
        G.ox = h_misc[0] - pad - grid_skin; G.oy = h_misc[1] - pad - grid_skin; G.oz = h_misc[2] - pad - grid_skin;
        G.hx = h_misc[3] + pad + grid_skin; G.hy = h_misc[4] + pad + grid_skin; G.hz = h_misc[5] + pad + grid_skin;
        const long long cell_cap = std::max<long long>(1 << 22, 8 * n_sup);
        for (;;) {
            G.nx = (int)std::floor((G.hx - G.ox) / G.edge) + 1;
            G.ny = (int)std::floor((G.hy - G.oy) / G.edge) + 1;
            G.nz = (int)std::floor((G.hz - G.oz) / G.edge) + 1;
            n_cell = (long long)G.nx * G.ny * G.nz;
            if (n_cell <= cell_cap) break;
            G.edge *= 1.26;     // very sparse configuration: coarser cells stay correct
        }
        const double par[7] = {G.ox, G.oy, G.oz, G.hx, G.hy, G.hz, G.edge};
        memcpy(nl->grid_par, par, sizeof par);
        nl->grid_dim[0] = G.nx; nl->grid_dim[1] = G.ny; nl->grid_dim[2] = G.nz;
        memcpy(nl->grid_box, h_misc, sizeof nl->grid_box);
        nl->grid_n = n_atoms;
        nl->grid_rsearch = basis->tab.r_search;
        nl->grid_cfirst = nl->c_first;
        nl->grid_ccount = nl->c_count;
        if (!is_device_pointer(image_offsets)) nl->grid_imgoff.assign(image_offsets, image_offsets + 3 * (size_t)n_images);
        else nl->grid_imgoff.clear();
    }
    BoxCheck chk = {};
    chk.active = reuse ? 1 : 0;
    chk.skin = grid_skin;
    for (int c = 0; c < 3; ++c) { chk.lo[c] = nl->grid_box[c]; chk.hi[c] = nl->grid_box[3 + c]; }
    UF3B_CUDA(nl->cell_of.reserve((size_t)n_sup));
    UF3B_CUDA(nl->cell_start.reserve((size_t)n_cell + 1));
    UF3B_CUDA(nl->cell_cursor.reserve((size_t)n_cell));
    UF3B_CUDA(cudaMemsetAsync(nl->cell_start.p, 0, sizeof(int) * (n_cell + 1), stream));
    UF3B_CUDA(cudaMemsetAsync(nl->cell_cursor.p, 0, sizeof(int) * n_cell, stream));
    const dim3 bin_grid((unsigned)std::min(std::max((n + 255) / 256, 1), 148 * 4), (unsigned)std::min(n_images, 65535));
    UF3B_LAUNCH(k_bin_count, bin_grid, 256, 0, stream, n, n_images, nl->pos.p, nl->img_off.p, G, nl->misc.p,
                nl->cell_of.p, nl->cell_start.p);
    if (int rc = scan_arrays(nl, 1, nl->cell_start.p, nl->cell_start.p, nullptr, nullptr, (int)n_cell,
                             nl->totals.p, stream)) return rc;
    // every supercell atom that survives the box test has a slot; n_sup bounds it
    UF3B_CUDA(nl->slots.reserve((size_t)n_sup));
    UF3B_LAUNCH(k_bin_fill, bin_grid, 256, 0, stream, n, n_images, nl->pos.p, nl->img_off.p, G, nl->misc.p,
                nl->spec.p, nl->cell_of.p, nl->cell_start.p, nl->cell_cursor.p, nl->slots.p);

    // one pass; the index arrays keep their capacity from earlier builds (first guess below)
    int *status = (int *)(nl->totals.p + 1);        // 8 ints
    int *claims = (int *)(nl->totals.p + 8);        // 2 * NL_REGIONS ints
    int h_status[8];
    UF3B_CUDA(nl->cnt2.reserve((size_t)n + 1));
    UF3B_CUDA(nl->cnt3.reserve((size_t)n + 1));
    if (nl->idx2.cap == 0) UF3B_CUDA(nl->idx2.reserve((size_t)n * 80 + 1024));
    if (nl->idx3.cap == 0 && basis->tab.n_trios > 0) UF3B_CUDA(nl->idx3.reserve((size_t)n * 40 + 1024));
    if (nl->idx3.cap == 0) UF3B_CUDA(nl->idx3.reserve(64));
    for (int attempt = 0; attempt < 2; ++attempt) {
        UF3B_CUDA(nl->scratch2.reserve(nl->idx2.cap));
        UF3B_CUDA(nl->scratch3.reserve(nl->idx3.cap));
        UF3B_CUDA(cudaMemsetAsync(status, 0, 9 * sizeof(int), stream));
        UF3B_CUDA(cudaMemsetAsync(claims, 0, 2 * NL_REGIONS * sizeof(int), stream));
        if (nl->c_count < n) {      // rows of the atoms other ranks own stay empty
            UF3B_CUDA(cudaMemsetAsync(nl->cnt2.p, 0, sizeof(int) * n, stream));
            UF3B_CUDA(cudaMemsetAsync(nl->cnt3.p, 0, sizeof(int) * n, stream));
            UF3B_CUDA(cudaMemsetAsync(nl->off2.p, 0, sizeof(int) * n, stream));
            UF3B_CUDA(cudaMemsetAsync(nl->off3.p, 0, sizeof(int) * n, stream));
        }
        const int cap2 = (int)(std::min<size_t>(nl->idx2.cap, (size_t)INT32_MAX) / NL_REGIONS);
        const int cap3 = (int)(std::min<size_t>(nl->idx3.cap, (size_t)INT32_MAX) / NL_REGIONS);
        UF3B_LAUNCH(k_neighbors, (unsigned)n_cell, NL_WARPS * 32, 0, stream, basis->tab, G, nl->slots.p,
                    nl->cell_start.p, nl->c_first, nl->c_count, nl->off2.p, nl->cnt2.p, nl->off3.p, nl->cnt3.p, nl->idx2.p,
                    nl->idx3.p, nl->scratch2.p, nl->scratch3.p, cap2, cap3, status, claims);
        UF3B_LAUNCH(k_post_status, 1, 32, 0, stream, claims, status, (volatile int *)(nl->h_mapped + 8), nl->misc.p,
                    chk, nl->max3_hint);
        if (basis->deferred_lists && reuse && attempt == 0) {
            // MD steady state: no host wait here — the caller queues its kernels behind the list build
            // and the status is verified by nlist_resolve() while they run
            if (!nl->status_ev)
                UF3B_CUDA(cudaEventCreateWithFlags(&nl->status_ev, cudaEventDisableTiming
                                                   | (blocking_sync_enabled() ? cudaEventBlockingSync : 0)));
            UF3B_CUDA(cudaEventRecord(nl->status_ev, stream));
            nl->pending = true;
            nl->hint_used = false;
            nl->consumed_pending = false;
            nl->pending_stream = stream;
            nl->grid_valid = true;
            guard.armed = false;
            *inout = nl;
            return UF3B_OK;
        }
        UF3B_CUDA(stream_sync(stream));
        memcpy(h_status, nl->h_mapped + 8, sizeof h_status);
        if (reuse) {
            if (h_status[7] & 1) return fail(UF3B_ERR_ELEMENT, "configuration holds an element outside the basis");
            if (h_status[7] & 2) return fail(UF3B_ERR_INVALID, "non-finite position");
            if (h_status[6]) {      // the centres left the cached box: build again with a fresh grid
                guard.armed = false;
                *inout = nl;
                return uf3b_neighbors_build_range(basis, n_atoms, positions, atomic_numbers, n_images,
                                                  image_offsets, image_abc, first_centre, n_centres, inout, stream_);
            }
        }
        if (h_status[0] < 0 || h_status[1] < 0)
            return fail(UF3B_ERR_CAPACITY, "neighbour list exceeds int32 offsets");
        if (!h_status[2]) break;
        if (attempt == 1) return fail(UF3B_ERR_CAPACITY, "neighbour list did not fit after regrowth");
        // every region must hold the fullest one; reserve() adds 25 % headroom
        UF3B_CUDA(nl->idx2.reserve((size_t)(h_status[4] + 1) * NL_REGIONS));
        UF3B_CUDA(nl->idx3.reserve((size_t)(h_status[5] + 1) * NL_REGIONS));
    }
    nl->total2 = h_status[0];
    nl->total3 = h_status[1];
    nl->max3 = h_status[3];
    nl->max3_hint = nl->max3;
    nl->grid_valid = true;
    guard.armed = false;
    *inout = nl;
    return UF3B_OK;
}

int uf3b_neighbors_count(const uf3b_nlist *nl, int which, int64_t *n_entries) {
    if (!nl || !n_entries || (which != 2 && which != 3)) return fail(UF3B_ERR_INVALID, "bad argument");
    if (int rc = nlist_resolve(const_cast<uf3b_nlist *>(nl))) return rc;
    *n_entries = which == 2 ? nl->total2 : nl->total3;
    return UF3B_OK;
}

int uf3b_neighbors_export(const uf3b_nlist *nl, int which, int64_t *offsets, int64_t *supercell_index) {
    if (!nl || !offsets || (which != 2 && which != 3)) return fail(UF3B_ERR_INVALID, "bad argument");
    if (int rc = nlist_resolve(const_cast<uf3b_nlist *>(nl))) return rc;
    const int64_t total = which == 2 ? nl->total2 : nl->total3;
    const size_t n = (size_t)nl->n;
    std::vector<int> start(n), count(n), idx;
    UF3B_CUDA(cudaDeviceSynchronize());
    if (n) {
        UF3B_CUDA(cudaMemcpy(start.data(), which == 2 ? nl->off2.p : nl->off3.p, sizeof(int) * n, cudaMemcpyDeviceToHost));
        UF3B_CUDA(cudaMemcpy(count.data(), which == 2 ? nl->cnt2.p : nl->cnt3.p, sizeof(int) * n, cudaMemcpyDeviceToHost));
    }
    if (total) {
        if (!supercell_index) return fail(UF3B_ERR_INVALID, "null index buffer");
        size_t extent = 0;          // rows are spread over the claim regions of the index array
        for (size_t a = 0; a < n; ++a) extent = std::max(extent, (size_t)start[a] + (size_t)count[a]);
        idx.resize(extent);
        UF3B_CUDA(cudaMemcpy(idx.data(), which == 2 ? nl->idx2.p : nl->idx3.p, sizeof(int) * extent, cudaMemcpyDeviceToHost));
    }
    // rows live in claim order on the device; the export is the standard CSR by atom
    int64_t run = 0;
    for (size_t a = 0; a < n; ++a) {
        offsets[a] = run;
        for (int k = 0; k < count[a]; ++k) supercell_index[run + k] = idx[(size_t)start[a] + k];
        run += count[a];
    }
    offsets[n] = run;
    return UF3B_OK;
}

void uf3b_nlist_destroy(uf3b_nlist *nl) { delete nl; }

}  // extern "C"
