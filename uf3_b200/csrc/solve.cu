// solve.cu — the regularised normal-equation solve on the device (cuSOLVER).
//
// Replaces regression/least_squares.py:763-771 (lu_factorization: scipy lu_factor /
// lu_solve on (G + R^T R) c = b).  The matrix is p x p with p <= ~1000, so this is one
// getrf + getrs; it exists so that a fit whose Gram blocks were accumulated and
// all-reduced on the device never has to round-trip through host LAPACK.
#include <cusolverDn.h>

#include "common.cuh"

using namespace uf3b;

#define UF3B_SOLVER(expr)                                                                 \
    do {                                                                                  \
        cusolverStatus_t s_ = (expr);                                                     \
        if (s_ != CUSOLVER_STATUS_SUCCESS) {                                              \
            rc = fail(UF3B_ERR_CUDA, "%s: cusolver status %d", #expr, (int)s_);           \
            goto done;                                                                    \
        }                                                                                 \
    } while (0)
#define UF3B_CUDA_GOTO(expr)                                                              \
    do {                                                                                  \
        cudaError_t e_ = (expr);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            rc = fail(UF3B_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_));            \
            goto done;                                                                    \
        }                                                                                 \
    } while (0)

extern "C" int uf3b_solve(const double *a, const double *b, int32_t n, int32_t n_rhs, double *x,
                          void *stream_) {
    if (!a || !b || !x || n < 1 || n_rhs < 1) return fail(UF3B_ERR_INVALID, "bad argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = UF3B_OK;
    // one cuSOLVER handle per host thread and device, kept for the life of the process: creating
    // it costs tens of milliseconds, more than a p = 73 solve itself
    static thread_local cusolverDnHandle_t t_handle[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    cusolverDnHandle_t local = nullptr;
    cusolverDnHandle_t &handle = (dev >= 0 && dev < 16) ? t_handle[dev] : local;
    // device scratch of the solve, kept per host thread and device like the handle (grow-only; five
    // cudaMalloc / cudaFree pairs cost more than the factorisation of a 70 x 70 system)
    struct Scratch { void *p = nullptr; size_t cap = 0; };
    static thread_local Scratch t_scratch[16][5];
    Scratch local_scratch[5];
    Scratch *scratch = (dev >= 0 && dev < 16) ? t_scratch[dev] : local_scratch;
    auto reserve = [&](int k, size_t bytes, void **out) -> cudaError_t {
        if (bytes > scratch[k].cap) {
            if (scratch[k].p) cudaFree(scratch[k].p);
            scratch[k].p = nullptr;
            scratch[k].cap = 0;
            cudaError_t e = cudaMalloc(&scratch[k].p, bytes + bytes / 4);
            if (e != cudaSuccess) return e;
            scratch[k].cap = bytes + bytes / 4;
        }
        *out = scratch[k].p;
        return cudaSuccess;
    };
    double *d_a = nullptr, *d_b = nullptr, *d_work = nullptr;
    int *d_piv = nullptr, *d_info = nullptr;
    int lwork = 0, h_info = 0;
    const size_t a_bytes = sizeof(double) * (size_t)n * n, b_bytes = sizeof(double) * (size_t)n * n_rhs;
    UF3B_CUDA_GOTO(reserve(0, a_bytes, (void **)&d_a));
    UF3B_CUDA_GOTO(reserve(1, b_bytes, (void **)&d_b));
    UF3B_CUDA_GOTO(reserve(2, sizeof(int) * n, (void **)&d_piv));
    UF3B_CUDA_GOTO(reserve(3, sizeof(int), (void **)&d_info));
    // `a` is row-major: LAPACK sees its transpose, so the solve below uses op = T
    UF3B_CUDA_GOTO(cudaMemcpyAsync(d_a, a, a_bytes, cudaMemcpyDefault, stream));
    UF3B_CUDA_GOTO(cudaMemcpyAsync(d_b, b, b_bytes, cudaMemcpyDefault, stream));   // [n_rhs][n]: column-major n x n_rhs
    if (!handle) UF3B_SOLVER(cusolverDnCreate(&handle));
    UF3B_SOLVER(cusolverDnSetStream(handle, stream));
    UF3B_SOLVER(cusolverDnDgetrf_bufferSize(handle, n, n, d_a, n, &lwork));
    UF3B_CUDA_GOTO(reserve(4, sizeof(double) * (size_t)std::max(lwork, 1), (void **)&d_work));
    UF3B_SOLVER(cusolverDnDgetrf(handle, n, n, d_a, n, d_work, d_piv, d_info));
    // the factorisation's status is read before getrs overwrites devInfo with its own: a zero pivot
    // (all-zero regulariser with uncovered columns) must not come back as inf / NaN coefficients
    UF3B_CUDA_GOTO(cudaMemcpyAsync(&h_info, d_info, sizeof(int), cudaMemcpyDeviceToHost, stream));
    UF3B_CUDA_GOTO(uf3b::stream_sync(stream));
    if (h_info != 0) {
        rc = fail(UF3B_ERR_STATE, "singular normal equations (getrf info %d)", h_info);
        goto done;
    }
    UF3B_SOLVER(cusolverDnDgetrs(handle, CUBLAS_OP_T, n, n_rhs, d_a, n, d_piv, d_b, n, d_info));
    UF3B_CUDA_GOTO(cudaMemcpyAsync(&h_info, d_info, sizeof(int), cudaMemcpyDeviceToHost, stream));
    UF3B_CUDA_GOTO(cudaMemcpyAsync(x, d_b, b_bytes, cudaMemcpyDefault, stream));
    UF3B_CUDA_GOTO(uf3b::stream_sync(stream));
    if (h_info != 0) rc = fail(UF3B_ERR_INVALID, "getrs rejected its arguments (info %d)", h_info);
done:
    if (local) cusolverDnDestroy(local);
    if (scratch == local_scratch)
        for (int k = 0; k < 5; ++k) cudaFree(local_scratch[k].p);
    return rc;
}
