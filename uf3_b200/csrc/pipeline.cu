// pipeline.cu — frames through host buffers, several in flight.
//
// The reference hands one frame at a time to `evaluate_configuration` (process.py:121-174) and
// fits from the stored rows afterwards (least_squares.py:355-433).  Here a caller streaming
// frames keeps `depth` of them in flight: every slot owns a basis handle (its scratch buffers),
// a neighbour-list handle, a stream and a worker thread that runs the C-ABI calls of the fit
// path for the slot's frame with HOST input pointers.  Two kinds of job:
//   rows  (uf3b_pipeline_submit)      uf3b_neighbors_build + uf3b_featurize into the caller's
//                                     host arrays — the rows of a frame cross PCIe (17.5 MB for
//                                     10 000 atoms and 73 columns);
//   fit   (uf3b_pipeline_submit_fit)  the same two calls with the rows left in HBM, then
//                                     uf3b_gram_accumulate into the slot's own normal-equation
//                                     accumulator: only positions / targets go up and the energy
//                                     row (F doubles) comes back.  uf3b_pipeline_export_gram sums
//                                     the slots' accumulators.
// A list build ends with a host synchronisation; with one thread per slot those waits only
// stall their own slot, so the device always has the other slots' kernels and copies queued.
#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

using namespace uf3b;

namespace {

struct Job {
    int64_t n = 0;
    const double *pos = nullptr;
    const int32_t *num = nullptr;
    std::vector<double> img_off;
    std::vector<int32_t> img_abc;
    double *xe = nullptr, *xf = nullptr;
    int64_t ld = 0;
    bool fit = false;
    const double *y = nullptr;      // fit: host force targets [3n], or null (energy row only)
};

struct PipeSlot {
    uf3b_basis *basis = nullptr;
    uf3b_nlist *nl = nullptr;
    uf3b_gram *gram = nullptr;      // fit jobs
    DevBuf<double> rows, d_y, d_xe;
    double moments[3] = {0.0, 0.0, 0.0};    // force targets seen: count, sum, sum of squares
    cudaStream_t stream = nullptr;
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    Job job;
    bool has_job = false, running = false, stop = false;
    int64_t ticket = -1;          // of the job last accepted
    int rc = UF3B_OK;
    std::string err;
};

int run_fit(PipeSlot *s, const Job &job) {
    const int F = s->basis->n_feats;
    const int64_t n = job.n;
    int rc = uf3b_neighbors_build(s->basis, n, job.pos, job.num, (int32_t)(job.img_off.size() / 3),
                                  job.img_off.data(), job.img_abc.data(), &s->nl, s->stream);
    if (rc != UF3B_OK) return rc;
    const bool forces = job.y != nullptr && n > 0;
    UF3B_CUDA(s->d_xe.reserve((size_t)F));
    if (forces) {
        UF3B_CUDA(s->rows.reserve((size_t)3 * n * F));
        UF3B_CUDA(s->d_y.reserve((size_t)3 * n));
        UF3B_CUDA(cudaMemcpyAsync(s->d_y.p, job.y, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, s->stream));
    }
    // everything below is queued without a host wait; the one synchronisation at the end also
    // covers the energy-row copy
    rc = uf3b_featurize(s->basis, s->nl, job.xe ? s->d_xe.p : nullptr, forces ? s->rows.p : nullptr, F, s->stream);
    if (rc != UF3B_OK) return rc;
    if (forces) {
        if (!s->gram) {
            rc = uf3b_gram_create(F, &s->gram);
            if (rc != UF3B_OK) return rc;
        }
        rc = uf3b_gram_accumulate(s->gram, s->rows.p, s->d_y.p, 3 * n, F, 1, s->stream);
        if (rc != UF3B_OK) return rc;
        double sum = 0.0, sq = 0.0;         // host side of the targets' statistics, overlapped with the kernels
        for (int64_t k = 0; k < 3 * n; ++k) { sum += job.y[k]; sq += job.y[k] * job.y[k]; }
        s->moments[0] += (double)(3 * n);
        s->moments[1] += sum;
        s->moments[2] += sq;
    }
    if (job.xe) UF3B_CUDA(cudaMemcpyAsync(job.xe, s->d_xe.p, sizeof(double) * F, cudaMemcpyDeviceToHost, s->stream));
    UF3B_CUDA(stream_sync(s->stream));
    return UF3B_OK;
}

void work(PipeSlot *s, int device) {
    cudaSetDevice(device);
    for (;;) {
        Job job;
        {
            std::unique_lock<std::mutex> lk(s->m);
            s->cv.wait(lk, [&] { return s->has_job || s->stop; });
            if (s->stop) return;
            job = std::move(s->job);
            s->has_job = false;
        }
        int rc;
        if (job.fit) {
            rc = run_fit(s, job);
        } else {
            rc = uf3b_neighbors_build(s->basis, job.n, job.pos, job.num, (int32_t)(job.img_off.size() / 3),
                                      job.img_off.data(), job.img_abc.data(), &s->nl, s->stream);
            if (rc == UF3B_OK) rc = uf3b_featurize(s->basis, s->nl, job.xe, job.xf, job.ld, s->stream);
        }
        {
            std::lock_guard<std::mutex> lk(s->m);
            s->rc = rc;
            s->err = rc == UF3B_OK ? "" : uf3b_last_error();
            s->running = false;
        }
        s->cv.notify_all();
    }
}

}  // namespace

struct uf3b_pipeline {
    int device = 0;
    int n_feats = 0;
    int64_t next = 0;
    std::vector<PipeSlot *> slots;
    // first failure of a frame nobody waited for before its slot was reused: kept so that it is
    // not lost (reported by the next submit / wait / export)
    int sticky_rc = UF3B_OK;
    std::string sticky_err;
};

namespace {

// Wait until the slot's previous frame is out and queue `job` on it.
int enqueue(uf3b_pipeline *p, Job &&job, int64_t *ticket) {
    if (p->sticky_rc != UF3B_OK) return fail(p->sticky_rc, "an earlier frame failed: %s", p->sticky_err.c_str());
    PipeSlot *s = p->slots[(size_t)(p->next % (int64_t)p->slots.size())];
    {
        std::unique_lock<std::mutex> lk(s->m);
        s->cv.wait(lk, [&] { return !s->running && !s->has_job; });
        if (s->rc != UF3B_OK) {             // the frame this slot held failed and was never waited for
            p->sticky_rc = s->rc;
            p->sticky_err = s->err;
            s->rc = UF3B_OK;
            return fail(p->sticky_rc, "an earlier frame failed: %s", p->sticky_err.c_str());
        }
        s->job = std::move(job);
        s->has_job = true;
        s->running = true;
        s->ticket = p->next;
    }
    s->cv.notify_all();
    *ticket = p->next++;
    return UF3B_OK;
}

int check_frame(const uf3b_pipeline *p, int64_t n_atoms, const double *positions, const int32_t *numbers,
                int32_t n_images, const double *image_offsets, const int32_t *image_abc, const int64_t *ticket) {
    if (!p || !ticket) return fail(UF3B_ERR_INVALID, "null pipeline / ticket");
    if (n_atoms < 0) return fail(UF3B_ERR_INVALID, "negative n_atoms");
    if (n_atoms > 0 && (!positions || !numbers)) return fail(UF3B_ERR_INVALID, "null positions / atomic numbers");
    if (n_images < 1 || !image_offsets || !image_abc) return fail(UF3B_ERR_INVALID, "bad image table");
    return UF3B_OK;
}

}  // namespace

extern "C" {

void uf3b_pipeline_destroy(uf3b_pipeline *p) {
    if (!p) return;
    for (PipeSlot *s : p->slots) {
        if (s->th.joinable()) {
            {
                std::unique_lock<std::mutex> lk(s->m);
                s->cv.wait(lk, [&] { return !s->running && !s->has_job; });
                s->stop = true;
            }
            s->cv.notify_all();
            s->th.join();
        }
        if (s->nl) uf3b_nlist_destroy(s->nl);
        if (s->gram) uf3b_gram_destroy(s->gram);
        if (s->basis) uf3b_basis_destroy(s->basis);
        if (s->stream) cudaStreamDestroy(s->stream);
        delete s;
    }
    delete p;
}

int uf3b_pipeline_create(const uf3b_basis_desc *desc, int32_t depth, uf3b_pipeline **out) {
    if (!desc || !out || depth < 1 || depth > 16) return fail(UF3B_ERR_INVALID, "bad argument (depth 1..16)");
    uf3b_pipeline *p = new uf3b_pipeline();
    cudaGetDevice(&p->device);
    p->n_feats = desc->n_feats;
    for (int k = 0; k < depth; ++k) {
        PipeSlot *s = new PipeSlot();
        p->slots.push_back(s);
        int rc = uf3b_basis_create(desc, &s->basis);
        if (rc == UF3B_OK && cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess)
            rc = fail(UF3B_ERR_CUDA, "cudaStreamCreate failed");
        if (rc != UF3B_OK) {
            uf3b_pipeline_destroy(p);
            return rc;
        }
        s->th = std::thread(work, s, p->device);
    }
    *out = p;
    return UF3B_OK;
}

int uf3b_pipeline_submit(uf3b_pipeline *p, int64_t n_atoms, const double *positions,
                         const int32_t *atomic_numbers, int32_t n_images, const double *image_offsets,
                         const int32_t *image_abc, double *x_energy, double *x_forces, int64_t ld,
                         int64_t *ticket) {
    if (int rc = check_frame(p, n_atoms, positions, atomic_numbers, n_images, image_offsets, image_abc, ticket)) return rc;
    if (x_forces && ld < p->n_feats) return fail(UF3B_ERR_INVALID, "ld smaller than n_feats");
    Job job;
    job.n = n_atoms;
    job.pos = positions;
    job.num = atomic_numbers;
    job.img_off.assign(image_offsets, image_offsets + 3 * (size_t)n_images);
    job.img_abc.assign(image_abc, image_abc + 3 * (size_t)n_images);
    job.xe = x_energy;
    job.xf = x_forces;
    job.ld = ld;
    return enqueue(p, std::move(job), ticket);
}

int uf3b_pipeline_submit_fit(uf3b_pipeline *p, int64_t n_atoms, const double *positions,
                             const int32_t *atomic_numbers, int32_t n_images, const double *image_offsets,
                             const int32_t *image_abc, const double *y_forces, double *x_energy,
                             int64_t *ticket) {
    if (int rc = check_frame(p, n_atoms, positions, atomic_numbers, n_images, image_offsets, image_abc, ticket)) return rc;
    Job job;
    job.n = n_atoms;
    job.pos = positions;
    job.num = atomic_numbers;
    job.img_off.assign(image_offsets, image_offsets + 3 * (size_t)n_images);
    job.img_abc.assign(image_abc, image_abc + 3 * (size_t)n_images);
    job.xe = x_energy;
    job.y = y_forces;
    job.fit = true;
    return enqueue(p, std::move(job), ticket);
}

int uf3b_pipeline_wait(uf3b_pipeline *p, int64_t ticket) {
    if (!p || ticket < 0 || ticket >= p->next) return fail(UF3B_ERR_INVALID, "unknown ticket");
    PipeSlot *s = p->slots[(size_t)(ticket % (int64_t)p->slots.size())];
    std::unique_lock<std::mutex> lk(s->m);
    if (s->ticket != ticket) {
        if (p->sticky_rc != UF3B_OK) return fail(p->sticky_rc, "an earlier frame failed: %s", p->sticky_err.c_str());
        return fail(UF3B_ERR_STATE, "the slot of this ticket has been reused");
    }
    s->cv.wait(lk, [&] { return !s->running; });
    if (s->rc != UF3B_OK) {
        const int rc = s->rc;
        s->rc = UF3B_OK;                    // reported: the slot may be reused
        return fail(rc, "%s", s->err.c_str());
    }
    return UF3B_OK;
}

int uf3b_pipeline_export_gram(uf3b_pipeline *p, double *gram_out, double *ord_out, double *moments_out) {
    if (!p) return fail(UF3B_ERR_INVALID, "null pipeline");
    const size_t F = (size_t)p->n_feats;
    if (gram_out) std::fill(gram_out, gram_out + F * F, 0.0);
    if (ord_out) std::fill(ord_out, ord_out + F, 0.0);
    if (moments_out) moments_out[0] = moments_out[1] = moments_out[2] = 0.0;
    std::vector<double> g(gram_out ? F * F : 0), b(ord_out ? F : 0);
    for (PipeSlot *s : p->slots) {
        std::unique_lock<std::mutex> lk(s->m);
        s->cv.wait(lk, [&] { return !s->running && !s->has_job; });
        if (s->rc != UF3B_OK) return fail(s->rc, "a frame failed: %s", s->err.c_str());
        if (moments_out)
            for (int k = 0; k < 3; ++k) moments_out[k] += s->moments[k];
        if (!s->gram) continue;
        if (int rc = uf3b_gram_export(s->gram, 1, gram_out ? g.data() : nullptr, ord_out ? b.data() : nullptr)) return rc;
        if (gram_out) for (size_t k = 0; k < F * F; ++k) gram_out[k] += g[k];
        if (ord_out) for (size_t k = 0; k < F; ++k) ord_out[k] += b[k];
    }
    if (p->sticky_rc != UF3B_OK) return fail(p->sticky_rc, "an earlier frame failed: %s", p->sticky_err.c_str());
    return UF3B_OK;
}

}  // extern "C"
