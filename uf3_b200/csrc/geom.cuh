// geom.cuh — device helpers shared by the neighbour, featurize and evaluate kernels.
#pragma once
#include "common.cuh"

namespace uf3b {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int pair_index(int ne, int a, int b) {
    if (a > b) { int t = a; a = b; b = t; }
    return a * ne - a * (a - 1) / 2 + (b - a);
}

struct Vec3 { double x, y, z; };

// image rank g = m / n of supercell index m = g * n + atom (0 <= m < 2^31), exact: the
// multiply-high estimate is at most one too small, one correction step fixes it.
__device__ __forceinline__ int image_of(const FrameView &f, int m) {
    int g = (int)__umulhi((unsigned)m, f.n_magic);
    if (m - g * f.n >= f.n) ++g;
    return g;
}

// Ghost position exactly as the reference builds it (data/geometry.py:146-147):
// positions + image offset, one rounded addition per component.
__device__ __forceinline__ Vec3 super_position(const FrameView &f, int m, int &atom) {
    const int g = image_of(f, m);
    atom = m - g * f.n;
    Vec3 p;
    p.x = __dadd_rn(__ldg(f.pos + 3 * atom + 0), __ldg(f.img_off + 3 * g + 0));
    p.y = __dadd_rn(__ldg(f.pos + 3 * atom + 1), __ldg(f.img_off + 3 * g + 1));
    p.z = __dadd_rn(__ldg(f.pos + 3 * atom + 2), __ldg(f.img_off + 3 * g + 2));
    return p;
}

__device__ __forceinline__ Vec3 real_position(const FrameView &f, int atom) {
    Vec3 p;
    p.x = __ldg(f.pos + 3 * atom + 0);
    p.y = __ldg(f.pos + 3 * atom + 1);
    p.z = __ldg(f.pos + 3 * atom + 2);
    return p;
}

// Euclidean distance with scipy cdist's operation order and no FMA contraction, so the
// strict / inclusive cutoff tests decide exactly like the reference's dense masks
// (representation/distances.py:66,134; angles.py:340,502-507).
__device__ __forceinline__ double dist_rn(const Vec3 &p, const Vec3 &q) {
    const double dx = __dsub_rn(p.x, q.x), dy = __dsub_rn(p.y, q.y), dz = __dsub_rn(p.z, q.z);
    return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
}

// 1/d for a finite, normal d > 0 (an interatomic distance): single-precision seed and two
// Newton steps, relative error < 1e-15 — used for direction cosines only, never for a
// cutoff decision.
__device__ __forceinline__ double fast_rcp(double d) {
    double r = (double)__frcp_rn((float)d);
    r = r * (2.0 - d * r);
    r = r * (2.0 - d * r);
    return r;
}

// ---- shared-memory access by 32-bit shared-window address.  The featurize kernel keeps
// its per-warp bases in registers through `pin` (ptxas otherwise rematerialises the
// generic->shared conversion, ~12 instructions, at every use in the hot loop); the
// accessors are volatile so that read-modify-write sequences keep their program order.
__device__ __forceinline__ unsigned smem_addr(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ unsigned pin(unsigned v) {
    asm volatile("mov.u32 %0, %0;" : "+r"(v));
    return v;
}
__device__ __forceinline__ double lds64(unsigned a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds32(unsigned a) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts64(unsigned a, double v) {
    asm volatile("st.shared.f64 [%0], %1;" :: "r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ double2 lds128(unsigned a) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ int4 lds128i(unsigned a) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(unsigned a, double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(a), "d"(v.x), "d"(v.y) : "memory");
}

// ---- TMA bulk copy (cp.async.bulk, completion counted on an mbarrier)
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}

// prior generic-proxy accesses of shared memory ordered before a following bulk copy into it
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(FULL, v, s);
    return v;
}

}  // namespace uf3b
