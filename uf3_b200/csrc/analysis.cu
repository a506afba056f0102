// analysis.cu — pair-distance histograms over the neighbour list.
//
// Replaces the counting part of distances.summarize_distances
// (representation/distances.py:367-442): there a dense (atoms x supercell) distance
// matrix is masked per pair interaction and handed to np.histogram; here the pair list of
// Kernel A (built with bounds (0, r_cut) for every pair) is walked once, one warp per
// centre, and every entry increments its (pair, bin) counter.  Bin lookup reproduces
// np.histogram on uniform edges: the truncated quotient, corrected against the actual edges.
#include <algorithm>

#include "common.cuh"
#include "geom.cuh"

namespace uf3b {

__global__ void __launch_bounds__(256)
k_pair_histogram(const BasisTab B, const FrameView f, const double *__restrict__ edges, int n_bins,
                 unsigned long long *__restrict__ hist) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_gw = (gridDim.x * blockDim.x) >> 5;
    const double first = edges[0], last = edges[n_bins];
    const double norm = (double)n_bins / (last - first);
    for (int a = f.c_first + gw; a < f.c_first + f.c_count; a += n_gw) {
        const int sa = __ldg(f.spec + a);
        const Vec3 pa = real_position(f, a);
        const int r0 = __ldg(f.off2 + a), r1 = r0 + __ldg(f.cnt2 + a);
        for (int e = r0 + lane; e < r1; e += 32) {
            int aj;
            const Vec3 pj = super_position(f, __ldg(f.idx2 + e), aj);
            const double d = dist_rn(pa, pj);
            if (!(d >= first && d <= last)) continue;
            int i = (int)((d - first) * norm);
            if (i >= n_bins) i = n_bins - 1;
            if (d < edges[i]) --i;
            else if (i != n_bins - 1 && d >= edges[i + 1]) ++i;
            const int pr = pair_index(B.ne, sa, __ldg(f.spec + aj));
            atomicAdd(hist + (size_t)pr * n_bins + i, 1ull);
        }
    }
}

}  // namespace uf3b

using namespace uf3b;

extern "C" int uf3b_pair_histogram(uf3b_basis *basis, const uf3b_nlist *nl, const double *bin_edges,
                                   int32_t n_bins, int64_t *counts, void *stream_) {
    if (!basis || !nl || !bin_edges || !counts || n_bins < 1) return fail(UF3B_ERR_INVALID, "bad argument");
    DeviceGuard on_device(basis->device);
    if (int rc = nlist_resolve(const_cast<uf3b_nlist *>(nl))) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int n_pairs = basis->tab.n_pairs;
    const size_t n_out = (size_t)n_pairs * n_bins;
    DevBuf<unsigned long long> d_hist;
    DevBuf<double> d_edges;
    UF3B_CUDA(d_hist.reserve(n_out));
    UF3B_CUDA(d_edges.reserve((size_t)n_bins + 1));
    UF3B_CUDA(cudaMemsetAsync(d_hist.p, 0, sizeof(unsigned long long) * n_out, stream));
    UF3B_CUDA(cudaMemcpyAsync(d_edges.p, bin_edges, sizeof(double) * (n_bins + 1), cudaMemcpyDefault, stream));
    if (nl->n > 0 && nl->c_count > 0) {
        const int blocks = std::max(1, std::min((nl->c_count + 7) / 8, sm_count() * 8));
        UF3B_LAUNCH(k_pair_histogram, blocks, 256, 0, stream, basis->tab, nl->view(), d_edges.p, n_bins, d_hist.p);
    }
    UF3B_CUDA(cudaMemcpyAsync(counts, d_hist.p, sizeof(long long) * n_out, cudaMemcpyDefault, stream));
    UF3B_CUDA(stream_sync(stream));
    return UF3B_OK;
}
