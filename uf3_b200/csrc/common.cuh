// common.cuh — shared declarations of libuf3b (sm_100a): error handling, device buffers,
// the device-side basis table and the per-configuration frame / neighbour-list views.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/uf3b.h"

struct uf3b_nlist;
struct uf3b_gram;

namespace uf3b {

// ------------------------------------------------------------------ host utilities
int fail(int code, const char *fmt, ...);
extern std::atomic<long long> g_launches;
extern std::atomic<bool> g_timing;            // read and written by the pipeline's worker threads
extern std::atomic<double> g_last_kernel_ms;

#define UF3B_CUDA(expr)                                                                   \
    do {                                                                                  \
        cudaError_t e_ = (expr);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return uf3b::fail(UF3B_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_));    \
    } while (0)

// kernel<<<grid, block, smem, stream>>>(args...) with launch accounting and error check.
#define UF3B_LAUNCH(kernel, grid, block, smem, stream, ...)                               \
    do {                                                                                  \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                       \
        uf3b::g_launches.fetch_add(1, std::memory_order_relaxed);                         \
        cudaError_t e_ = cudaGetLastError();                                              \
        if (e_ != cudaSuccess)                                                            \
            return uf3b::fail(UF3B_ERR_CUDA, "launch %s: %s", #kernel,                    \
                              cudaGetErrorString(e_));                                    \
    } while (0)

// Grow-only device buffer owned by a handle.
template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        size_t want = n + n / 4 + 16;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
};

inline bool is_device_pointer(const void *ptr) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

int sm_count();      // of the current device

// Raises a kernel's dynamic shared-memory limit to at least `bytes` and never lowers it: handles
// with different bases (and the pipeline's worker threads) share the kernels, so a per-call exact
// value could shrink the limit between another thread's call and its launch.
cudaError_t ensure_dynamic_smem(const void *kernel, size_t bytes);

// Makes the handle's device current for the duration of a C-ABI call and restores the caller's.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

// Host wait for `stream`.  Default: cudaStreamSynchronize (spins: lowest latency).  With
// uf3b_set_blocking_sync(1) the thread sleeps on an event created with cudaEventBlockingSync
// instead — for hosts where ranks x pipeline workers outnumber the cores.
cudaError_t stream_sync(cudaStream_t stream);
bool blocking_sync_enabled();

// ------------------------------------------------------------------ device tables
// Flattened BSplineBasis on the device.  Pair p = (a<=b) -> a*ne - a*(a-1)/2 + (b-a);
// trio t = centre*n_pairs + pair(j,k).  Every knot vector owns (n_knots - 7) cubic
// pieces of 16 doubles: piece i-3 holds, for the four basis functions that are
// non-zero on (t[i], t[i+1]], the coefficients of u^0..u^3 with u = r - t[i].
struct BasisTab {
    int ne, n_pairs, n_trios, n_feats;
    int lead2, trail2, lead3, trail3;
    double r3min, r3max;      // angles.py:312-325
    double r_search;          // max over pair r_max and r3max (cell edge)
    const int *pair_nk, *pair_koff, *pair_poff, *pair_col;
    const double *pair_lo, *pair_hi;   // strict bounds max(r_min,0) < d < r_max
    const double *knots2, *poly2, *pair_scale;   // scale: intervals per unit length
    const int *trio_nk, *trio_koff, *trio_poff;   // [3*t + leg]
    const int *trio_col, *trio_goff, *trio_sym;   // [t]
    const double *knots3, *poly3, *trio_scale;   // [3*t + leg]
    const int *bin_col;       // full grid bin -> compressed column (or -1), trio-major
    const double *bin_w;      // folding weight of the bin
    const double *coeff;      // flat model coefficients [n_feats]   (inference only)
    const double *c_grid;     // decompressed 3-body coefficient grids, bin layout
    const int *z_to_spec;     // [128] atomic number -> element index or -1
    int unit_weights;         // every kept bin folds with weight exactly 1 (always true for
                              // tables produced from compress_3B; lets kernels skip bin_w)
};

// One binned supercell atom (32 B so that a run of cells is one aligned bulk copy).
struct __align__(32) Slot {
    double x, y, z;
    int m;        // supercell index  image_rank * n_atoms + atom
    int spec;     // element index
};

// Read-only view of one configuration and its neighbour lists.
struct FrameView {
    int n;                    // real atoms
    unsigned n_magic;         // floor(2^32 / n) (2^32 - 1 for n = 1): image_of() without a division
    int n_img;
    int c_first, c_count;     // centres whose rows were built (all atoms unless the frame is
                              // split over ranks: uf3b_neighbors_build_range)
    const double *pos;        // [n*3]
    const int *spec;          // [n]
    const double *img_off;    // [n_img*3]
    const int *img_inv;       // [n_img] rank of the image with negated coordinates
    // per-centre rows (start, count) into the index arrays; every row sorted by supercell index
    const int *off2, *cnt2, *idx2;   // pair list   (max(r_min,0) < d < r_max per pair)
    const int *off3, *cnt3, *idx3;   // 3-body list (r3min < d <= r3max)
};

}  // namespace uf3b

namespace uf3b {
// Waits for the deferred status of a list build and validates it: UF3B_OK, an error code, or
// UF3B_RETRY when the build has to be repeated (the centres left the cached grid, or the index
// arrays were too small — they have been regrown).  No-op for a list that is not pending.
int nlist_resolve(uf3b_nlist *nl);
// uf3b_gram_accumulate whose kernels leave at once when *invalid != 0 (device flag of a deferred list build)
int gram_accumulate_guarded(uf3b_gram *gram, const double *x, const double *y, int64_t rows, int64_t ld,
                            int is_force, void *stream, const int *invalid);
int gram_clear(uf3b_gram *gram, cudaStream_t stream);                           // both accumulators = 0
int gram_add(uf3b_gram *dst, const uf3b_gram *src, cudaStream_t stream);        // dst += src
}  // namespace uf3b

// ------------------------------------------------------------------ opaque handles
struct uf3b_basis {
    uf3b::BasisTab tab;            // device pointers into `blob`
    void *blob = nullptr;          // one allocation holding every table
    double *coeff = nullptr;       // [n_feats]
    double *c_grid = nullptr;      // [n_bins]
    int n_bins = 0;
    int n_feats = 0;
    int n_knots2 = 0, n_poly2 = 0, n_knots3 = 0, n_poly3 = 0;   // doubles in the spline tables
    bool has_coeff = false;
    int device = 0;
    std::vector<int> h_trio_goff, h_bin_col, h_trio_col, h_trio_sym, h_trio_dims;   // dims: [3*t + leg]
    std::vector<int> h_trio_koff, h_trio_poff;     // [3*t + leg] offsets (doubles) into knots3 / poly3
    std::vector<double> h_trio_scale;              // [3*t + leg]
    std::vector<double> h_knots3;                  // host copy of the 3-body knot vectors
    int h_pair_nk0 = 0;                            // knots of pair 0
    bool no_tile = false;          // force the general scatter path (tests / profiling)
    bool deferred_lists = false;   // list builds that reuse their grid return without the host wait
    int frames_in_flight = 1;      // k_featurize launches take 1/k of the resident blocks
    std::vector<double> h_bin_w;
    std::vector<int> h_numbers;
    // scratch for energy partial sums and force-row staging (grow-only)
    uf3b::DevBuf<double> partials;
    uf3b::DevBuf<double> stage;
    uf3b::DevBuf<double> stage_e;
    uf3b::DevBuf<double> gacc;     // global-memory accumulators for very wide rows
    uf3b::DevBuf<unsigned char> leg_cache;   // k_leg_cache records of the current frame
    uf3b::DevBuf<double> legv, legd, epos;   // k_centre_legs tables of the current frame (tiled path)
    uf3b::DevBuf<double> planes;             // plane table of the current frame (tiled path)
    uf3b::DevBuf<double> tile3;              // neighbour-role force tiles handed from k_rows_nbr to k_rows_ctr
};

struct uf3b_nlist {
    int64_t n = 0;
    int n_img = 0;
    int64_t total2 = 0, total3 = 0;
    int max3 = 0;                  // longest row of the 3-body list
    int c_first = 0, c_count = 0;  // centre range of the last build
    // 16 doubles of pinned, device-mapped host memory: the two small read-backs of a build
    // (bounding box, list totals) are WRITTEN there by a kernel instead of being copied, so
    // they never queue behind a large device->host row copy on the copy engine
    double *h_mapped = nullptr;
    // deferred status (uf3b_basis_set_deferred_lists): the build returned without waiting for its
    // totals / overflow / box check; nlist_resolve() waits on `status_ev` and validates
    bool pending = false;
    int max3_hint = 0;             // longest 3-body row of the last verified build
    bool hint_used = false;        // a consumer sized itself by max3_hint while the build was pending
    bool consumed_pending = false; // feature rows were produced from a build that is still pending
    bool ticket_zeroed = false;    // k_prepare's block counter has been cleared once
    cudaEvent_t status_ev = nullptr;
    cudaStream_t pending_stream = nullptr;
    ~uf3b_nlist() {
        if (h_mapped) cudaFreeHost(h_mapped);
        if (status_ev) cudaEventDestroy(status_ev);
    }
    // cell grid of the last build, reusable while the centres stay within `grid_skin` of the
    // box it was sized for (MD: one host synchronisation per build instead of two)
    bool grid_valid = false;
    double grid_par[7] = {0, 0, 0, 0, 0, 0, 0};   // ox oy oz hx hy hz edge
    int grid_dim[3] = {0, 0, 0};
    double grid_box[6] = {0, 0, 0, 0, 0, 0};      // bounding box of the centres it was built for
    long long grid_n = -1;
    double grid_rsearch = 0.0;
    int grid_cfirst = 0, grid_ccount = 0;
    std::vector<double> grid_imgoff;
    std::vector<double> img_host;                  // image table that is on the device (img_off / img_inv) ...
    std::vector<int32_t> abc_host;
    int img_uploaded = 0;                          // ... and its length, 0 = none
    uf3b::DevBuf<double> pos, img_off;
    uf3b::DevBuf<int> z, spec, img_inv;
    uf3b::DevBuf<int> off2, off3, cnt2, cnt3, idx2, idx3, scratch2, scratch3;
    uf3b::DevBuf<int> cell_of, cell_start, cell_cursor;
    uf3b::DevBuf<uf3b::Slot> slots;
    uf3b::DevBuf<double> misc;     // bbox (6) on device
    uf3b::DevBuf<long long> totals, tile_sums;
    uf3b::FrameView view() const;
    // device flag (1 = a consumer queued behind a deferred build must skip the frame), or null
    const int *invalid_flag() const { return totals.p ? reinterpret_cast<const int *>(totals.p + 1) + 8 : nullptr; }
};
