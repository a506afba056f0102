/*
 * uf3b.h — C ABI of the B200-native UF3 hot path (libuf3b.so, sm_100a).
 *
 * The reference (uf3/uf3, pure Python) has no FFI for this path: its boundary is the
 * Python object API.  This header is the boundary a binding would
 * target instead; each entry point names the reference code it replaces.  The Python
 * host layer in uf3_b200/ (process.BasisFeaturizer, calculator.UFCalculator) calls
 * exactly these functions through ctypes — see INTEGRATION.md for the stub a reference
 * maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative uf3b_status otherwise; the
 *     message for the calling thread's last failure is uf3b_last_error()
 *   - handles are opaque; one handle per (process, device); not thread-safe
 *   - the caller owns every buffer passed in; array arguments may be HOST or DEVICE
 *     pointers (detected with cudaPointerGetAttributes) — device pointers avoid the
 *     host round trip in frame loops
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls with
 *     host output buffers synchronise that stream before returning
 *   - all floating point is IEEE float64; indices are int32 on device, int64 on export
 *   - "supercell index" = image_rank * n_atoms + atom, image_rank in the order of the
 *     image table handed to uf3b_neighbors_build (reference: data/geometry.py:141-149)
 */
#ifndef UF3B_H
#define UF3B_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UF3B_ABI_VERSION 1

typedef enum {
    UF3B_OK = 0,
    UF3B_ERR_INVALID = -1,      /* bad argument / inconsistent descriptor          */
    UF3B_ERR_CUDA = -2,         /* CUDA runtime error (message has the cudaError)  */
    UF3B_ERR_ELEMENT = -3,      /* atomic number not in the basis' element list    */
    UF3B_ERR_CAPACITY = -4,     /* index range exceeded (int32 offsets)            */
    UF3B_ERR_STATE = -5,        /* call order violated (e.g. coefficients not set) */
    UF3B_RETRY = 1              /* deferred list build (uf3b_basis_set_deferred_lists) was invalid: build the
                                   lists again and re-issue the call; nothing was written that may be used */
} uf3b_status;

typedef struct uf3b_basis uf3b_basis;     /* device-resident basis tables            */
typedef struct uf3b_nlist uf3b_nlist;     /* one configuration: binned ghosts + CSR  */
typedef struct uf3b_gram uf3b_gram;       /* normal-equation accumulator             */

/*
 * Flattened BSplineBasis (reference: representation/bspline.py:20-720, the host-side
 * descriptor that stays the user-facing API).  Interactions are indexed the way
 * ChemicalSystem orders them (data/composition.py:113-132): elements by ascending Z;
 * pair (a<=b) -> a*ne - a*(a-1)/2 + (b-a); trio (c; a<=b) -> c*n_pairs + pair(a,b).
 * Feature columns follow bspline.py:525-575: [one per element][pairs][trios], WITHOUT
 * the leading "y" column.
 */
typedef struct {
    int32_t n_elements;
    const int32_t *atomic_numbers;   /* [n_elements], ascending                                 */
    int32_t n_feats;                 /* total feature columns                                   */
    int32_t leading_trim_2b, trailing_trim_2b;   /* bspline.py:66-67, applied per basis index  */
    int32_t leading_trim_3b, trailing_trim_3b;
    /* pairs */
    const int32_t *pair_n_knots;     /* [n_pairs]                                               */
    const double *pair_knots;        /* concatenated knot sequences                             */
    const double *pair_r_min;        /* [n_pairs] r_min_map; bound used is max(r_min, 0)        */
    const double *pair_r_max;        /* [n_pairs] r_max_map                                     */
    const int32_t *pair_col;         /* [n_pairs] first feature column                          */
    /* trios (n_trios = 0 for a 2-body basis, else n_elements * n_pairs) */
    int32_t n_trios;
    const int32_t *trio_n_knots;     /* [n_trios*3] l, m, n legs                                */
    const double *trio_knots;        /* concatenated, trio-major then leg                       */
    const int32_t *trio_col;         /* [n_trios] first feature column                          */
    const int32_t *trio_n_cols;      /* [n_trios] compressed column count                       */
    /* compression (bspline.py:664-719 as a map): for every bin of the full L*M*N grid,  */
    /* trio-major and C-ordered: compressed column (or -1) and folding weight.             */
    const int32_t *bin_col;
    const double *bin_weight;
    const int32_t *trio_symmetry;    /* [n_trios] 1, 2 or 3 (bspline.py:723-763)                */
} uf3b_basis_desc;

const char *uf3b_last_error(void);
int uf3b_abi_version(void);

/* Device used by subsequent create calls (default: current device). */
int uf3b_set_device(int device);
/* Number of CUDA devices this process can see (BasisFeaturizer.evaluate_parallel deals its
 * batches over them; the reference deals them over worker processes, process.py:196-254). */
int uf3b_device_count(int32_t *count);
/* Host waits inside the library sleep on a blocking event instead of spinning in
 * cudaStreamSynchronize (process-wide; for hosts where ranks x pipeline workers outnumber cores). */
int uf3b_set_blocking_sync(int enabled);

/* -- basis ------------------------------------------------------------------------ */
/* replaces: BSplineBasis.update_basis_functions / generate_basis_functions
 * (bspline.py:322-369, :791-807): builds per-interval cubic pieces of every basis
 * function and uploads all tables. */
int uf3b_basis_create(const uf3b_basis_desc *desc, uf3b_basis **out);
/* replaces: calculator.coefficients_by_interaction / construct_pair_potentials /
 * construct_trio_potentials (forcefield/calculator.py:490-573).  `coefficients` is the
 * flat model vector of length n_feats (host pointer). */
int uf3b_basis_set_coefficients(uf3b_basis *basis, const double *coefficients, int32_t n);
/* Throughput knob for streams of frames (no counterpart in the reference): with k handles
 * working on k consecutive frames on k streams, let every uf3b_featurize launch occupy only
 * 1/k of the SM resources, so that the frames in flight share each SM.  Default 1. */
int uf3b_basis_set_frames_in_flight(uf3b_basis *basis, int32_t k);
/* MD loops: a list build on this basis that can reuse the cell grid of the previous one (same frame
 * size, atoms still inside the padded box) returns WITHOUT waiting for the device — its totals /
 * overflow flag / box check are verified by the next call that uses the list, after that call has
 * queued its own kernels.  uf3b_energy_forces then returns UF3B_RETRY (1) if the build was invalid.
 * uf3b_featurize with DEVICE outputs does not wait either (a device-side flag makes its kernels leave at
 * once behind an invalid build); the rows it produced are verified by the handle's NEXT
 * uf3b_neighbors_build*, which returns UF3B_RETRY when they must not be used (repeat that frame, then
 * the build), and by uf3b_neighbors_count / uf3b_neighbors_export for the last frame of a stream. */
int uf3b_basis_set_deferred_lists(uf3b_basis *basis, int enabled);
void uf3b_basis_destroy(uf3b_basis *basis);

/* -- neighbour lists -------------------------------------------------------------- */
/* replaces: geometry.get_supercell + distances.get_distance_matrix + the boolean masks
 * (data/geometry.py:14-51; representation/distances.py:19-75,146-169;
 * representation/angles.py:289-346).
 *   positions      [n_atoms*3]  row-major x,y,z
 *   atomic_numbers [n_atoms]
 *   image_offsets  [n_images*3] cartesian offset of every periodic image, image 0 = 0
 *   image_abc      [n_images*3] integer image coordinates (to pair image g with -g)
 * Builds, for every real atom, (a) the pair list  max(r_min,0) < d < r_max  per
 * interaction and (b) the 3-body list  r3_min < d <= r3_max, both sorted by supercell
 * index.  If *inout is non-NULL its buffers are reused. */
int uf3b_neighbors_build(uf3b_basis *basis, int64_t n_atoms, const double *positions,
                         const int32_t *atomic_numbers, int32_t n_images,
                         const double *image_offsets, const int32_t *image_abc,
                         uf3b_nlist **inout, void *stream);
/* Same for the centres [first_centre, first_centre + n_centres) only (the other rows stay
 * empty): the atom-range partition of ONE large frame over ranks (MD strong scaling).  All
 * atoms are still binned, so every rank sees every neighbour.  uf3b_energy_forces on such a
 * list returns this rank's PARTIAL energy, forces [n_atoms*3] (own atoms' pair and centre
 * terms plus the 3-body reactions on any atom) and virial; the sum over ranks is the total.
 * uf3b_featurize refuses a partial list. */
int uf3b_neighbors_build_range(uf3b_basis *basis, int64_t n_atoms, const double *positions,
                               const int32_t *atomic_numbers, int32_t n_images,
                               const double *image_offsets, const int32_t *image_abc,
                               int64_t first_centre, int64_t n_centres,
                               uf3b_nlist **inout, void *stream);
/* Number of entries in list `which` (2 or 3). */
int uf3b_neighbors_count(const uf3b_nlist *nl, int which, int64_t *n_entries);
/* Parity hook: CSR offsets [n_atoms+1] and supercell indices [n_entries] (host). */
int uf3b_neighbors_export(const uf3b_nlist *nl, int which, int64_t *offsets,
                          int64_t *supercell_index);
void uf3b_nlist_destroy(uf3b_nlist *nl);

/* -- fit path --------------------------------------------------------------------- */
/* replaces: BasisFeaturizer.featurize_energy_2B/3B, featurize_force_2B/3B and the row
 * assembly of evaluate_configuration (representation/process.py:293-506).
 *   x_energy [n_feats] or NULL           (columns incl. the element counts)
 *   x_forces [3*n_atoms rows] or NULL    row c*n_atoms + a, `ld` doubles apart
 *                                        (ld >= n_feats; ld = n_feats+1 leaves room
 *                                        for a y column when the pointer is offset) */
int uf3b_featurize(uf3b_basis *basis, const uf3b_nlist *nl, double *x_energy,
                   double *x_forces, int64_t ld, void *stream);

/* -- inference path --------------------------------------------------------------- */
/* replaces: UFCalculator._get_potential_energy / _get_forces
 * (forcefield/calculator.py:156-343).  energy: 1 double or NULL; forces [n_atoms*3] or
 * NULL; virial [9] (row-major 3x3, symmetric) or NULL: W = dE/d(strain) = sum over every leg
 * of every pair / triplet term of (dE/dr) r u (x) u; stress = W / volume.  Replaces the
 * twelve strained-cell energy evaluations of calculator.py:399-404. */
int uf3b_energy_forces(uf3b_basis *basis, const uf3b_nlist *nl, double *energy,
                       double *forces, double *virial, void *stream);

/* -- normal equations (regression/least_squares.py:733-771) ------------------------ */
/* G += X^T X, b += X^T y over `rows` rows of X (row stride ld, n_cols columns, device
 * or host).  Two accumulators are kept, selected by is_force. */
int uf3b_gram_create(int32_t n_cols, uf3b_gram **out);
int uf3b_gram_accumulate(uf3b_gram *gram, const double *x, const double *y, int64_t rows,
                         int64_t ld, int is_force, void *stream);
int uf3b_gram_export(const uf3b_gram *gram, int is_force, double *gram_out, double *ord_out);
void uf3b_gram_destroy(uf3b_gram *gram);
/* -- frames in flight through host buffers --------------------------------------------- */
/* `depth` slots, each with its own handles, stream and worker thread; a submitted frame runs
 * uf3b_neighbors_build + uf3b_featurize with the HOST pointers given (pinned memory makes the
 * row copy asynchronous to the other slots' kernels).  The reference processes one
 * configuration at a time (process.py:121-174).  positions / numbers / x_energy / x_forces must
 * stay valid until uf3b_pipeline_wait(ticket) returns; a slot is reused every `depth`
 * submissions (submit blocks while its previous frame is still running).  Submit and wait from
 * one thread. */
typedef struct uf3b_pipeline uf3b_pipeline;
int uf3b_pipeline_create(const uf3b_basis_desc *desc, int32_t depth, uf3b_pipeline **out);
int uf3b_pipeline_submit(uf3b_pipeline *pipe, int64_t n_atoms, const double *positions,
                         const int32_t *atomic_numbers, int32_t n_images, const double *image_offsets,
                         const int32_t *image_abc, double *x_energy, double *x_forces, int64_t ld,
                         int64_t *ticket);
int uf3b_pipeline_wait(uf3b_pipeline *pipe, int64_t ticket);
/* Fit job: the same frame through uf3b_neighbors_build + uf3b_featurize with the 3N x F force rows
 * LEFT IN HBM, then uf3b_gram_accumulate into the slot's own normal-equation accumulator — what
 * BasisFeaturizer.batched_to_hdf + WeightedLinearModel.fit_from_file do through an HDF5 file
 * (process.py:256-291, least_squares.py:355-433).  y_forces: host targets [3N] in row order
 * (fx_0.., fy_0.., fz_0..) or NULL for an energy-only frame; x_energy: host array [F] that receives
 * the frame's energy row (or NULL).  Only positions and targets go up; F doubles come back. */
int uf3b_pipeline_submit_fit(uf3b_pipeline *pipe, int64_t n_atoms, const double *positions,
                             const int32_t *atomic_numbers, int32_t n_images, const double *image_offsets,
                             const int32_t *image_abc, const double *y_forces, double *x_energy,
                             int64_t *ticket);
/* Waits for every slot, then sums the slots' force accumulators: gram_out [F*F] row-major, ord_out [F],
 * moments_out [3] = number of force targets, their sum and sum of squares (any may be NULL). */
int uf3b_pipeline_export_gram(uf3b_pipeline *pipe, double *gram_out, double *ord_out, double *moments_out);
void uf3b_pipeline_destroy(uf3b_pipeline *pipe);

/* -- analysis ----------------------------------------------------------------------- */
/* Pair-distance histogram per pair interaction over list 2 of `nl` (replaces the counting of
 * distances.summarize_distances, representation/distances.py:401-423): bin_edges [n_bins+1]
 * uniform and ascending, counts [n_pairs*n_bins] (pair-major, the basis' pair order); every
 * (real centre, neighbour) entry is counted, as the reference's masked distance matrix does. */
int uf3b_pair_histogram(uf3b_basis *basis, const uf3b_nlist *nl, const double *bin_edges,
                        int32_t n_bins, int64_t *counts, void *stream);

/* Dense solve A x = b on the device (cuSOLVER getrf + getrs), the regularised normal
 * equations of regression/least_squares.py:248-272,763-771.  a [n*n] row-major, b and x
 * [n_rhs][n]; host or device pointers; synchronises the stream. */
int uf3b_solve(const double *a, const double *b, int32_t n, int32_t n_rhs, double *x, void *stream);

/* -- host-side probe ----------------------------------------------------------------- */
/* Runs the table builder of uf3b_basis_create on one knot vector and evaluates the four
 * non-zero cubic basis functions and their first derivatives at r ON THE HOST (no GPU
 * needed): the same polynomial pieces and the same evaluation routine the kernels use.
 * Returns the first basis index (searchsorted(knots, r, 'left') - 4, bspline.py:966),
 * or -1 when r is outside (knots[3], knots[n-4]].  Used by the CPU test-suite. */
int uf3b_host_eval_basis(const double *knots, int32_t n_knots, double r, double *v, double *dv);

/* -- instrumentation ---------------------------------------------------------------- */
/* Kernels launched by this library in this process since load (for bench.py). */
int64_t uf3b_launch_count(void);
/* Device time (ms) of the most recent uf3b_featurize / uf3b_energy_forces main kernel,
 * measured with CUDA events on the call's stream; enable with uf3b_set_timing(1). */
int uf3b_set_timing(int enabled);
double uf3b_last_kernel_ms(void);
/* Measured non-tensor float64 FMA rate of the current device (dependent DFMA chains):
 * the second roof the 3-body kernels are reported against in bench.py. */
int uf3b_probe_fp64_tflops(double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* UF3B_H */
