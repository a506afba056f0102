// pipeline.cu — frames through host buffers, several in flight.
//
// The reference hands one frame at a time to `evaluate_configuration` (process.py:121-174) and
// fits from the stored rows afterwards (least_squares.py:355-433).  Here a caller streaming
// frames keeps `depth` of them in flight: every slot owns a basis handle (its scratch buffers), a
// neighbour-list handle and a stream; the pipeline's thread queues a frame's whole chain on the
// slot's stream without waiting for the device (the list build of a frame reuses the cell grid of
// the slot's previous frame and is verified afterwards: uf3b_basis_set_deferred_lists).  Two jobs:
//   rows  (uf3b_pipeline_submit)      lists + feature rows, copied into the caller's host arrays —
//                                     the rows of a frame cross PCIe (17.5 MB for 10 000 atoms and
//                                     73 columns);
//   fit   (uf3b_pipeline_submit_fit)  the rows stay in HBM and go straight into the slot's own
//                                     normal-equation accumulator (uf3b_gram_accumulate): only
//                                     positions / targets go up and the energy row (F doubles) comes
//                                     back.  uf3b_pipeline_export_gram sums the slots' accumulators.
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <deque>
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

enum SlotState { IDLE, QUEUED, INFLIGHT, DONE };

struct PipeSlot {
    uf3b_basis *basis = nullptr;
    uf3b_nlist *nl = nullptr;
    uf3b_gram *gram = nullptr;      // fit jobs
    DevBuf<double> rows, d_y, d_xe;
    double moments[3] = {0.0, 0.0, 0.0};    // force targets seen: count, sum, sum of squares
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    Job job;
    SlotState state = IDLE;
    int64_t ticket = -1;          // of the job last accepted
    int rc = UF3B_OK;
    std::string err;
    long long n_jobs = 0, n_retries = 0;
};

}  // namespace

struct uf3b_pipeline {
    int device = 0;
    int n_feats = 0;
    int64_t next = 0;
    std::vector<PipeSlot *> slots;
    // ONE thread issues every CUDA call of the pipeline: frames are queued on their slot's stream
    // without host waits (deferred list builds) and completed — event wait, verification of the
    // list build, repetition of the rare invalid frame — in submission order.  (One thread per slot
    // was measured first: the threads spinning in cudaStreamSynchronize and launching at the same
    // time cost 0.4-0.8 ms of queueing per frame against 0.05 ms from a single thread.)
    std::thread worker;
    std::mutex m;
    std::condition_variable cv;
    std::deque<int64_t> queue, inflight;
    bool stop = false;
    // first failure of a frame nobody waited for before its slot was reused: kept so that it is
    // not lost (reported by the next submit / wait / export)
    int sticky_rc = UF3B_OK;
    std::string sticky_err;
    uf3b_gram *total = nullptr;     // sum of the slots' normal equations (uf3b_pipeline_export_gram)
};

namespace {

// The frame's chain on the slot's stream.  sync = true: checked list build and a wait at the end
// (first frame of a slot, repetition of an invalid frame).
int issue(uf3b_pipeline *p, PipeSlot *s, bool sync) {
    const Job &job = s->job;
    const int F = p->n_feats;
    const int64_t n = job.n;
    uf3b_basis_set_deferred_lists(s->basis, sync ? 0 : 1);
    int rc = uf3b_neighbors_build(s->basis, n, job.pos, job.num, (int32_t)(job.img_off.size() / 3),
                                  job.img_off.data(), job.img_abc.data(), &s->nl, s->stream);
    if (rc != UF3B_OK) return rc;
    const bool forces = n > 0 && (job.fit ? job.y != nullptr : job.xf != nullptr);
    UF3B_CUDA(s->d_xe.reserve((size_t)F));
    if (forces) UF3B_CUDA(s->rows.reserve((size_t)3 * n * F));
    if (job.fit && forces) {
        UF3B_CUDA(s->d_y.reserve((size_t)3 * n));
        UF3B_CUDA(cudaMemcpyAsync(s->d_y.p, job.y, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, s->stream));
    }
    rc = uf3b_featurize(s->basis, s->nl, job.xe ? s->d_xe.p : nullptr, forces ? s->rows.p : nullptr, F, s->stream);
    if (rc != UF3B_OK) return rc;
    if (job.fit && forces) {
        if (!s->gram) {
            rc = uf3b_gram_create(F, &s->gram);
            if (rc != UF3B_OK) return rc;
        }
        // rows behind a deferred build that turns out invalid must not reach the normal equations
        rc = gram_accumulate_guarded(s->gram, s->rows.p, s->d_y.p, 3 * n, F, 1, s->stream,
                                     s->nl->pending ? s->nl->invalid_flag() : nullptr);
        if (rc != UF3B_OK) return rc;
    }
    if (!job.fit && forces) {
        if (job.ld == F)
            UF3B_CUDA(cudaMemcpyAsync(job.xf, s->rows.p, sizeof(double) * F * (size_t)3 * n, cudaMemcpyDeviceToHost, s->stream));
        else
            UF3B_CUDA(cudaMemcpy2DAsync(job.xf, sizeof(double) * job.ld, s->rows.p, sizeof(double) * F, sizeof(double) * F,
                                        (size_t)3 * n, cudaMemcpyDeviceToHost, s->stream));
    }
    if (job.xe) UF3B_CUDA(cudaMemcpyAsync(job.xe, s->d_xe.p, sizeof(double) * F, cudaMemcpyDeviceToHost, s->stream));
    UF3B_CUDA(cudaEventRecord(s->done, s->stream));
    if (sync) UF3B_CUDA(cudaEventSynchronize(s->done));
    return UF3B_OK;
}

// Wait for the slot's frame and verify the list build it ran behind; repeat the frame if needed.
int complete(uf3b_pipeline *p, PipeSlot *s) {
    if (s->rc != UF3B_OK) return s->rc;
    UF3B_CUDA(cudaEventSynchronize(s->done));
    int rc = nlist_resolve(s->nl);
    if (rc == UF3B_RETRY) {
        s->n_retries++;
        rc = issue(p, s, true);
    }
    if (rc == UF3B_OK && s->job.fit && s->job.y && s->job.n > 0) {
        double sum = 0.0, sq = 0.0;
        for (int64_t k = 0; k < 3 * s->job.n; ++k) { sum += s->job.y[k]; sq += s->job.y[k] * s->job.y[k]; }
        s->moments[0] += (double)(3 * s->job.n);
        s->moments[1] += sum;
        s->moments[2] += sq;
    }
    return rc;
}

void work(uf3b_pipeline *p) {
    cudaSetDevice(p->device);
    std::unique_lock<std::mutex> lk(p->m);
    for (;;) {
        p->cv.wait(lk, [&] { return p->stop || !p->queue.empty() || !p->inflight.empty(); });
        if (!p->queue.empty()) {
            const int64_t t = p->queue.front();
            p->queue.pop_front();
            PipeSlot *s = p->slots[(size_t)(t % (int64_t)p->slots.size())];
            lk.unlock();
            // the first frame of a slot has no grid to reuse: its build is a checked one anyway
            int rc = issue(p, s, false);
            std::string err = rc == UF3B_OK ? "" : uf3b_last_error();
            lk.lock();
            s->rc = rc;
            s->err = err;
            s->state = INFLIGHT;
            s->n_jobs++;
            p->inflight.push_back(t);
        } else if (!p->inflight.empty()) {
            const int64_t t = p->inflight.front();
            PipeSlot *s = p->slots[(size_t)(t % (int64_t)p->slots.size())];
            lk.unlock();
            int rc = complete(p, s);
            std::string err = rc == UF3B_OK ? "" : (s->rc != UF3B_OK ? s->err : std::string(uf3b_last_error()));
            lk.lock();
            s->rc = rc;
            s->err = err;
            s->state = DONE;
            p->inflight.pop_front();
            p->cv.notify_all();
        } else {
            return;         // stop requested and nothing left
        }
    }
}

// Wait until the slot's previous frame is out and queue `job` on it.
int enqueue(uf3b_pipeline *p, Job &&job, int64_t *ticket) {
    std::unique_lock<std::mutex> lk(p->m);
    if (p->sticky_rc != UF3B_OK) return fail(p->sticky_rc, "an earlier frame failed: %s", p->sticky_err.c_str());
    PipeSlot *s = p->slots[(size_t)(p->next % (int64_t)p->slots.size())];
    p->cv.wait(lk, [&] { return s->state == IDLE || s->state == DONE; });
    if (s->rc != UF3B_OK) {             // the frame this slot held failed and was never waited for
        p->sticky_rc = s->rc;
        p->sticky_err = s->err;
        s->rc = UF3B_OK;
        return fail(p->sticky_rc, "an earlier frame failed: %s", p->sticky_err.c_str());
    }
    s->job = std::move(job);
    s->state = QUEUED;
    s->ticket = p->next;
    p->queue.push_back(p->next);
    *ticket = p->next++;
    lk.unlock();
    p->cv.notify_all();
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
    if (p->worker.joinable()) {
        {
            std::lock_guard<std::mutex> lk(p->m);
            p->stop = true;
        }
        p->cv.notify_all();
        p->worker.join();           // drains the queue and the frames in flight first
    }
    for (PipeSlot *s : p->slots) {
        if (getenv("UF3B_PIPE_DEBUG") && s->n_jobs)
            fprintf(stderr, "[uf3b pipeline] slot: %lld frames, %lld repeated\n", s->n_jobs, s->n_retries);
        if (s->nl) uf3b_nlist_destroy(s->nl);
        if (s->gram) uf3b_gram_destroy(s->gram);
        if (s->basis) uf3b_basis_destroy(s->basis);
        if (s->done) cudaEventDestroy(s->done);
        if (s->stream) cudaStreamDestroy(s->stream);
        delete s;
    }
    if (p->total) uf3b_gram_destroy(p->total);
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
        // consecutive frames run on different slots' streams: each launch takes half of the SM resources, so
        // that the tail of one frame's kernels and the list build of the next share the chip
        // (uf3b_basis_set_frames_in_flight; UF3B_PIPE_IN_FLIGHT overrides)
        if (rc == UF3B_OK && depth >= 2) {
            const char *env = getenv("UF3B_PIPE_IN_FLIGHT");
            rc = uf3b_basis_set_frames_in_flight(s->basis, env ? std::max(1, atoi(env)) : 2);
        }
        if (rc == UF3B_OK && cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess)
            rc = fail(UF3B_ERR_CUDA, "cudaStreamCreate failed");
        if (rc == UF3B_OK && cudaEventCreateWithFlags(&s->done, cudaEventDisableTiming
                                                      | (blocking_sync_enabled() ? cudaEventBlockingSync : 0)) != cudaSuccess)
            rc = fail(UF3B_ERR_CUDA, "cudaEventCreate failed");
        if (rc != UF3B_OK) {
            uf3b_pipeline_destroy(p);
            return rc;
        }
    }
    p->worker = std::thread(work, p);
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
    std::unique_lock<std::mutex> lk(p->m);
    if (s->ticket != ticket) {
        if (p->sticky_rc != UF3B_OK) return fail(p->sticky_rc, "an earlier frame failed: %s", p->sticky_err.c_str());
        return fail(UF3B_ERR_STATE, "the slot of this ticket has been reused");
    }
    p->cv.wait(lk, [&] { return s->state == DONE || s->ticket != ticket; });
    if (s->ticket != ticket) return fail(UF3B_ERR_STATE, "the slot of this ticket has been reused");
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
    std::unique_lock<std::mutex> lk(p->m);
    p->cv.wait(lk, [&] {                                                            // every frame is out
        for (const PipeSlot *s : p->slots)
            if (s->state == QUEUED || s->state == INFLIGHT) return false;
        return true;
    });
    // every frame is complete (its slot's event was waited for): the slots' accumulators are summed on the
    // device and come back with ONE synchronisation and two copies
    DeviceGuard on_device(p->device);
    bool any = false;
    for (PipeSlot *s : p->slots) {
        if (s->rc != UF3B_OK) return fail(s->rc, "a frame failed: %s", s->err.c_str());
        if (moments_out)
            for (int k = 0; k < 3; ++k) moments_out[k] += s->moments[k];
        if (!s->gram) continue;
        if (!p->total)
            if (int rc = uf3b_gram_create((int32_t)F, &p->total)) return rc;
        if (!any)
            if (int rc = gram_clear(p->total, nullptr)) return rc;
        any = true;
        if (int rc = gram_add(p->total, s->gram, nullptr)) return rc;
    }
    if (any)
        if (int rc = uf3b_gram_export(p->total, 1, gram_out, ord_out)) return rc;
    if (p->sticky_rc != UF3B_OK) return fail(p->sticky_rc, "an earlier frame failed: %s", p->sticky_err.c_str());
    return UF3B_OK;
}

}  // extern "C"
