// basis.cu — library state, error reporting and the device-resident basis tables.
//
// uf3b_basis_create replaces BSplineBasis.update_basis_functions /
// generate_basis_functions (representation/bspline.py:322-369, :791-807);
// uf3b_basis_set_coefficients replaces coefficients_by_interaction /
// construct_pair_potentials / construct_trio_potentials
// (forcefield/calculator.py:490-573).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

#include "common.cuh"
#include "spline.cuh"

namespace uf3b {

static thread_local std::string t_error;
std::atomic<long long> g_launches{0};
std::atomic<bool> g_timing{false};
std::atomic<double> g_last_kernel_ms{0.0};

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_error = buf;
    return code;
}

static std::atomic<bool> g_blocking_sync{false};

bool blocking_sync_enabled() { return g_blocking_sync.load(std::memory_order_relaxed); }

cudaError_t stream_sync(cudaStream_t stream) {
    if (!g_blocking_sync.load(std::memory_order_relaxed)) return cudaStreamSynchronize(stream);
    thread_local cudaEvent_t ev = nullptr;
    if (!ev) {
        cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming);
        if (e != cudaSuccess) { ev = nullptr; return e; }
    }
    cudaError_t e = cudaEventRecord(ev, stream);
    return e != cudaSuccess ? e : cudaEventSynchronize(ev);
}

int sm_count() {
    static std::atomic<int> cached[64];        // per device
    int dev = 0;
    cudaGetDevice(&dev);
    const int slot = dev >= 0 && dev < 64 ? dev : 0;
    int n = cached[slot].load(std::memory_order_relaxed);
    if (n <= 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[slot].store(n, std::memory_order_relaxed);
    }
    return n;
}

cudaError_t ensure_dynamic_smem(const void *kernel, size_t bytes) {
    static std::mutex m;
    static std::map<std::pair<int, const void *>, size_t> limit;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(m);
    size_t &cur = limit[std::make_pair(dev, kernel)];
    if (bytes <= cur || bytes <= 48 * 1024) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) cur = bytes;
    return e;
}

namespace {

// Sequentially packs host arrays into one blob with 32-byte alignment.
struct Packer {
    std::vector<unsigned char> bytes;
    template <class T>
    size_t add(const std::vector<T> &v) {
        size_t off = (bytes.size() + 31) & ~size_t(31);
        bytes.resize(off + std::max<size_t>(v.size(), 1) * sizeof(T), 0);
        if (!v.empty()) memcpy(bytes.data() + off, v.data(), v.size() * sizeof(T));
        return off;
    }
};

}  // namespace
}  // namespace uf3b

using namespace uf3b;

extern "C" {

const char *uf3b_last_error(void) { return t_error.c_str(); }
int uf3b_abi_version(void) { return UF3B_ABI_VERSION; }
int64_t uf3b_launch_count(void) { return (int64_t)g_launches.load(); }
int uf3b_set_timing(int enabled) {
    g_timing.store(enabled != 0);
    return UF3B_OK;
}
double uf3b_last_kernel_ms(void) { return g_last_kernel_ms.load(); }

int uf3b_set_blocking_sync(int enabled) {
    g_blocking_sync.store(enabled != 0);
    return UF3B_OK;
}

int uf3b_set_device(int device) {
    UF3B_CUDA(cudaSetDevice(device));
    return UF3B_OK;
}

int uf3b_device_count(int32_t *count) {
    if (!count) return fail(UF3B_ERR_INVALID, "count is null");
    int n = 0;
    UF3B_CUDA(cudaGetDeviceCount(&n));
    *count = n;
    return UF3B_OK;
}

int uf3b_host_eval_basis(const double *knots, int32_t n_knots, double r, double *v, double *dv) {
    if (!knots || n_knots < 8 || !v || !dv) return fail(UF3B_ERR_INVALID, "bad knot vector");
    std::vector<double> poly;
    build_pieces(knots, n_knots, poly);
    return eval_leg(knots, n_knots, knot_scale(knots, n_knots), poly.data(), r, 0, 0, v, dv);
}

int uf3b_basis_create(const uf3b_basis_desc *d, uf3b_basis **out) {
    if (!d || !out) return fail(UF3B_ERR_INVALID, "null argument");
    const int ne = d->n_elements;
    if (ne < 1 || ne > 64) return fail(UF3B_ERR_INVALID, "n_elements out of range");
    const int n_pairs = ne * (ne + 1) / 2;
    const int n_trios = d->n_trios;
    if (n_trios != 0 && n_trios != ne * n_pairs)
        return fail(UF3B_ERR_INVALID, "n_trios must be 0 or n_elements * n_pairs");
    if (d->n_feats < ne) return fail(UF3B_ERR_INVALID, "n_feats smaller than n_elements");

    std::vector<int> z_to_spec(128, -1), numbers(ne);
    for (int e = 0; e < ne; ++e) {
        const int z = d->atomic_numbers[e];
        if (z < 0 || z >= 128) return fail(UF3B_ERR_INVALID, "atomic number out of range");
        if (e && z <= d->atomic_numbers[e - 1])
            return fail(UF3B_ERR_INVALID, "atomic_numbers must be ascending");
        z_to_spec[z] = e;
        numbers[e] = z;
    }

    // ---- pairs
    std::vector<int> pair_nk(n_pairs), pair_koff(n_pairs), pair_poff(n_pairs), pair_col(n_pairs);
    std::vector<double> pair_lo(n_pairs), pair_hi(n_pairs), knots2, poly2, pair_scale(n_pairs);
    double r_search = 0.0;
    {
        int koff = 0;
        for (int p = 0; p < n_pairs; ++p) {
            const int nk = d->pair_n_knots[p];
            if (nk < 8) return fail(UF3B_ERR_INVALID, "pair %d: fewer than 8 knots", p);
            pair_nk[p] = nk;
            pair_koff[p] = koff;
            pair_poff[p] = (int)poly2.size();
            for (int k = 0; k < nk; ++k) {
                const double t = d->pair_knots[koff + k];
                if (k && t < knots2.back()) return fail(UF3B_ERR_INVALID, "pair %d: knots not sorted", p);
                knots2.push_back(t);
            }
            build_pieces(d->pair_knots + koff, nk, poly2);
            pair_scale[p] = knot_scale(d->pair_knots + koff, nk);
            koff += nk;
            pair_lo[p] = d->pair_r_min[p] > 0.0 ? d->pair_r_min[p] : 0.0;   // distances.py:60
            pair_hi[p] = d->pair_r_max[p];
            pair_col[p] = d->pair_col[p];
            if (pair_col[p] < 0 || pair_col[p] + nk - 4 > d->n_feats)
                return fail(UF3B_ERR_INVALID, "pair %d: feature columns out of range", p);
            if (pair_hi[p] > r_search) r_search = pair_hi[p];
        }
    }

    // ---- trios
    std::vector<int> trio_nk(3 * n_trios), trio_koff(3 * n_trios), trio_poff(3 * n_trios);
    std::vector<int> trio_col(n_trios), trio_goff(n_trios), trio_sym(n_trios);
    std::vector<double> knots3, poly3, trio_scale(3 * n_trios);
    std::vector<int> bin_col;
    std::vector<double> bin_w;
    double r3min = 0.0, r3max = 0.0;
    int n_bins = 0;
    if (n_trios) {
        int koff = 0;
        double kmin = 0.0;
        bool first = true;
        for (int t = 0; t < n_trios; ++t) {
            long long grid = 1;
            for (int leg = 0; leg < 3; ++leg) {
                const int nk = d->trio_n_knots[3 * t + leg];
                if (nk < 8) return fail(UF3B_ERR_INVALID, "trio %d leg %d: fewer than 8 knots", t, leg);
                trio_nk[3 * t + leg] = nk;
                trio_koff[3 * t + leg] = koff;
                trio_poff[3 * t + leg] = (int)poly3.size();
                for (int k = 0; k < nk; ++k) {
                    const double v = d->trio_knots[koff + k];
                    if (k && v < knots3.back())
                        return fail(UF3B_ERR_INVALID, "trio %d leg %d: knots not sorted", t, leg);
                    knots3.push_back(v);
                    if (first || v < kmin) { kmin = v; first = false; }             // angles.py:312
                    if (leg < 2 && v > r3max) r3max = v;                            // angles.py:322
                }
                build_pieces(d->trio_knots + koff, nk, poly3);
                trio_scale[3 * t + leg] = knot_scale(d->trio_knots + koff, nk);
                koff += nk;
                grid *= nk - 4;
            }
            if ((long long)n_bins + grid > (1LL << 30)) return fail(UF3B_ERR_CAPACITY, "3-body grid too large");
            trio_goff[t] = n_bins;
            n_bins += (int)grid;
            trio_col[t] = d->trio_col[t];
            trio_sym[t] = d->trio_symmetry ? d->trio_symmetry[t] : 3;
            if (trio_col[t] < 0 || trio_col[t] + d->trio_n_cols[t] > d->n_feats)
                return fail(UF3B_ERR_INVALID, "trio %d: feature columns out of range", t);
        }
        r3min = kmin > 0.0 ? kmin : 0.0;                                            // angles.py:314
        bin_col.assign(d->bin_col, d->bin_col + n_bins);
        bin_w.assign(d->bin_weight, d->bin_weight + n_bins);
        for (int t = 0; t < n_trios; ++t) {
            const int end = (t + 1 < n_trios) ? trio_goff[t + 1] : n_bins;
            for (int b = trio_goff[t]; b < end; ++b)
                if (bin_col[b] >= d->trio_n_cols[t])
                    return fail(UF3B_ERR_INVALID, "trio %d: bin column out of range", t);
        }
        if (r3max > r_search) r_search = r3max;
    }
    if (!(r_search > 0.0)) r_search = 1.0;

    Packer pk;
    const size_t o_pair_nk = pk.add(pair_nk), o_pair_koff = pk.add(pair_koff);
    const size_t o_pair_poff = pk.add(pair_poff), o_pair_col = pk.add(pair_col);
    const size_t o_pair_lo = pk.add(pair_lo), o_pair_hi = pk.add(pair_hi);
    const size_t o_knots2 = pk.add(knots2), o_poly2 = pk.add(poly2);
    const size_t o_trio_nk = pk.add(trio_nk), o_trio_koff = pk.add(trio_koff);
    const size_t o_trio_poff = pk.add(trio_poff), o_trio_col = pk.add(trio_col);
    const size_t o_trio_goff = pk.add(trio_goff), o_trio_sym = pk.add(trio_sym);
    const size_t o_knots3 = pk.add(knots3), o_poly3 = pk.add(poly3);
    const size_t o_bin_col = pk.add(bin_col), o_bin_w = pk.add(bin_w);
    const size_t o_z = pk.add(z_to_spec);
    const size_t o_pair_scale = pk.add(pair_scale), o_trio_scale = pk.add(trio_scale);

    uf3b_basis *b = new uf3b_basis();
    cudaGetDevice(&b->device);
    cudaError_t e = cudaMalloc(&b->blob, pk.bytes.size());
    if (e == cudaSuccess) e = cudaMemcpy(b->blob, pk.bytes.data(), pk.bytes.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc((void **)&b->coeff, sizeof(double) * d->n_feats);
    if (e == cudaSuccess) e = cudaMalloc((void **)&b->c_grid, sizeof(double) * (n_bins > 0 ? n_bins : 1));
    if (e != cudaSuccess) {
        uf3b_basis_destroy(b);
        return fail(UF3B_ERR_CUDA, "basis upload: %s", cudaGetErrorString(e));
    }
    const unsigned char *base = (const unsigned char *)b->blob;
    BasisTab &T = b->tab;
    T.ne = ne; T.n_pairs = n_pairs; T.n_trios = n_trios; T.n_feats = d->n_feats;
    T.lead2 = d->leading_trim_2b; T.trail2 = d->trailing_trim_2b;
    T.lead3 = d->leading_trim_3b; T.trail3 = d->trailing_trim_3b;
    T.r3min = r3min; T.r3max = r3max; T.r_search = r_search;
    T.pair_nk = (const int *)(base + o_pair_nk);
    T.pair_koff = (const int *)(base + o_pair_koff);
    T.pair_poff = (const int *)(base + o_pair_poff);
    T.pair_col = (const int *)(base + o_pair_col);
    T.pair_lo = (const double *)(base + o_pair_lo);
    T.pair_hi = (const double *)(base + o_pair_hi);
    b->n_knots2 = (int)knots2.size(); b->n_poly2 = (int)poly2.size();
    b->n_knots3 = (int)knots3.size(); b->n_poly3 = (int)poly3.size();
    T.knots2 = (const double *)(base + o_knots2);
    T.poly2 = (const double *)(base + o_poly2);
    T.trio_nk = (const int *)(base + o_trio_nk);
    T.trio_koff = (const int *)(base + o_trio_koff);
    T.trio_poff = (const int *)(base + o_trio_poff);
    T.trio_col = (const int *)(base + o_trio_col);
    T.trio_goff = (const int *)(base + o_trio_goff);
    T.trio_sym = (const int *)(base + o_trio_sym);
    T.knots3 = (const double *)(base + o_knots3);
    T.poly3 = (const double *)(base + o_poly3);
    T.bin_col = (const int *)(base + o_bin_col);
    T.bin_w = (const double *)(base + o_bin_w);
    T.z_to_spec = (const int *)(base + o_z);
    T.pair_scale = (const double *)(base + o_pair_scale);
    T.trio_scale = (const double *)(base + o_trio_scale);
    T.unit_weights = 1;
    for (int k = 0; k < n_bins; ++k)
        if (bin_col[k] >= 0 && bin_w[k] != 1.0) T.unit_weights = 0;
    T.coeff = b->coeff;
    T.c_grid = b->c_grid;
    b->n_bins = n_bins;
    b->n_feats = d->n_feats;
    b->h_trio_goff = trio_goff;
    b->h_trio_col = trio_col;
    b->h_trio_sym = trio_sym;
    b->h_trio_dims.resize(3 * n_trios);
    for (int k = 0; k < 3 * n_trios; ++k) b->h_trio_dims[k] = trio_nk[k] - 4;
    b->h_trio_koff = trio_koff;
    b->h_trio_poff = trio_poff;
    b->h_trio_scale = trio_scale;
    b->h_knots3 = knots3;
    b->h_pair_nk0 = pair_nk[0];
    b->no_tile = getenv("UF3B_NO_TILE") != nullptr;
    b->h_bin_col = bin_col;
    b->h_bin_w = bin_w;
    b->h_numbers = numbers;
    *out = b;
    return UF3B_OK;
}

int uf3b_basis_set_frames_in_flight(uf3b_basis *b, int32_t k) {
    if (!b || k < 1 || k > 8) return fail(UF3B_ERR_INVALID, "frames in flight must be 1..8");
    b->frames_in_flight = k;
    return UF3B_OK;
}

int uf3b_basis_set_deferred_lists(uf3b_basis *b, int enabled) {
    if (!b) return fail(UF3B_ERR_INVALID, "null basis");
    b->deferred_lists = enabled != 0;
    return UF3B_OK;
}

int uf3b_basis_set_coefficients(uf3b_basis *b, const double *coefficients, int32_t n) {
    if (!b || !coefficients) return fail(UF3B_ERR_INVALID, "null argument");
    if (n != b->n_feats) return fail(UF3B_ERR_INVALID, "expected %d coefficients, got %d", b->n_feats, n);
    DeviceGuard on_device(b->device);
    // decompress_3B as a gather (bspline.py:693-719): grid[bin] = c[col(bin)] * w(bin)
    std::vector<double> grid(b->n_bins > 0 ? b->n_bins : 1, 0.0);
    const int n_trios = b->tab.n_trios;
    for (int t = 0; t < n_trios; ++t) {
        const int end = (t + 1 < n_trios) ? b->h_trio_goff[t + 1] : b->n_bins;
        for (int bin = b->h_trio_goff[t]; bin < end; ++bin) {
            const int col = b->h_bin_col[bin];
            if (col >= 0) grid[bin] = coefficients[b->h_trio_col[t] + col] * b->h_bin_w[bin];
        }
    }
    UF3B_CUDA(cudaMemcpy(b->coeff, coefficients, sizeof(double) * n, cudaMemcpyDefault));
    UF3B_CUDA(cudaMemcpy(b->c_grid, grid.data(), sizeof(double) * grid.size(), cudaMemcpyHostToDevice));
    b->has_coeff = true;
    return UF3B_OK;
}

void uf3b_basis_destroy(uf3b_basis *b) {
    if (!b) return;
    if (b->blob) cudaFree(b->blob);
    if (b->coeff) cudaFree(b->coeff);
    if (b->c_grid) cudaFree(b->c_grid);
    delete b;
}

}  // extern "C"
