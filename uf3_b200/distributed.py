"""Multi-GPU plumbing: frames shard over ranks, one all-reduce of the normal equations.

The reference parallelises featurization over frames with a process pool and merges
pandas DataFrames (`process.py:196-254`, `util/parallel.py:167-251`).  Frames are
independent, so here frame f goes to rank f mod world_size, every rank accumulates its
own Gram statistics (`least_squares.GramStats`, 2 F^2 + 2 F + 6 doubles) and a single
`all_reduce(SUM)` over NCCL (GPU ranks) or gloo (CPU tests) replaces `gather_and_merge`.
There is no collective on the featurization path itself.

Inference on ONE large frame (the MD loop of BASELINE.json configs[4]) strong-scales by an
atom-range partition instead: positions are replicated, rank r builds the neighbour rows
of its own contiguous atom range (`uf3b_neighbors_build_range`) and evaluates those centres;
the 3-body reactions it adds to atoms of other ranks make its force array a PARTIAL sum, so
one `all_reduce(SUM)` of 3N + 1 doubles (forces + energy; 2.4 MB at 100 000 atoms) per step
completes it (`ShardedEvaluator`).  The reference has no counterpart (its calculator is
single-process, forcefield/calculator.py:124-153).
"""
import os

import numpy as np


def shard(items, rank, world_size):
    """Round-robin shard of a sequence: rank r gets items r, r + world, ..."""
    return list(items)[rank::world_size]


def all_reduce_stats(stats, group=None):
    """Sum `GramStats` over all ranks in place (no-op without an initialised group)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats
    vec = torch.from_numpy(stats.to_vector())
    if dist.get_backend(group) == "nccl":
        vec = vec.cuda()
    dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    stats.from_vector(vec.cpu().numpy())
    return stats


def accumulate_frames(featurizer, frames, stats=None, rank=0, world_size=1):
    """Featurize this rank's share of `frames` on its GPU and fold the rows into Gram
    statistics without moving force rows off the device.

    frames: sequence of (geom, energy, forces) with forces shaped (3, N) or None."""
    import torch
    from uf3_b200.atoms import frame_arrays
    from uf3_b200 import geometry
    from uf3_b200.least_squares import GramAccumulator

    eng = featurizer.engine
    F = eng.n_feats
    if stats is None:
        stats = GramAccumulator(F)
    # every tensor lives on the ENGINE's device, and the kernels run on that device's current stream
    dev = torch.device("cuda", eng.device if getattr(eng, "device", None) is not None else torch.cuda.current_device())
    stream = torch.cuda.current_stream(dev).cuda_stream
    mine = shard(frames, rank, world_size)
    # energy rows stay on the device until the end (one small copy for the whole shard) and the
    # force targets go up asynchronously, so a frame costs no host synchronisation beyond the
    # one inside its list build
    xe_all = torch.zeros((max(len(mine), 1), F), dtype=torch.float64, device=dev)
    rows = None
    energy_rows = []
    for k, (geom, energy, forces) in enumerate(mine):
        positions, numbers, cell, pbc = frame_arrays(geom)
        images = geometry.image_table(cell, pbc, featurizer.r_cut) if np.any(pbc) else None
        eng.build_neighbors(positions, numbers, images=images, stream=stream)
        n = len(positions)
        want_f = forces is not None and featurizer.fit_forces and n > 0
        if want_f and (rows is None or rows.shape[0] < 3 * n):
            rows = torch.empty((3 * n, F), dtype=torch.float64, device=dev)
        eng.featurize_device(xe_all[k].data_ptr(), rows.data_ptr() if want_f else None, F, stream)
        if energy is not None:
            energy_rows.append((k, float(energy), n))
        if want_f:
            y = np.ascontiguousarray(forces, dtype=np.float64).reshape(-1)
            y_dev = torch.from_numpy(y).to(dev, non_blocking=True)
            stats.add_force_rows_device(rows.data_ptr(), y_dev.data_ptr(), 3 * n, F, stream,
                                        y_moments=(float(y.sum()), float(np.dot(y, y))))
            del y_dev        # freed by torch's caching allocator in stream order
    xe_host = xe_all.cpu().numpy()
    for k, energy, n in energy_rows:
        stats.add_energy_row(xe_host[k], energy, n)
    return stats


def accumulate_frames_pipelined(basis, frames, stats=None, rank=0, world_size=1, device=None, depth=4,
                                fit_forces=True):
    """This rank's share of `frames` through the library's own frame pipeline in FIT mode
    (`uf3b_pipeline_submit_fit`): host positions / targets in, force rows left in HBM and folded
    into the slots' normal-equation accumulators, one energy row (F doubles) back per frame.
    Nothing else crosses PCIe, so the throughput does not depend on the host's ingest bandwidth
    (the rows-to-host path is bound by it: 17.5 MB per 10 000-atom frame at 73 columns).

    frames: sequence of (geom, energy, forces) with forces shaped (3, N) or None; `geom` anything
    `uf3_b200.atoms.frame_arrays` understands.  Returns `GramStats` (host) holding this rank's
    sums; `all_reduce_stats` then makes them global with ONE collective — the replacement of
    `batched_to_hdf` + `fit_from_file` (process.py:256-291, least_squares.py:355-433)."""
    from uf3_b200 import geometry
    from uf3_b200.atoms import frame_arrays
    from uf3_b200.least_squares import GramStats
    from uf3_b200.pipeline import NativePipeline

    pipe = NativePipeline(basis, depth=depth, device=device)
    F = pipe.n_feats
    if stats is None:
        stats = GramStats(F)
    r_cut = basis.r_cut
    pending = []

    def finish(item):
        ticket, xe, energy, n = item
        pipe.wait(ticket)
        if energy is not None and n > 0:
            stats.add_energy_row(xe, energy, n)

    try:
        for geom, energy, forces in shard(frames, rank, world_size):
            positions, numbers, cell, pbc = frame_arrays(geom)
            images = geometry.image_table(cell, pbc, r_cut) if np.any(pbc) else \
                (np.zeros((1, 3), dtype=np.int64), np.zeros((1, 3)))
            y = None
            if forces is not None and fit_forces and len(positions) > 0:
                y = np.ascontiguousarray(forces, dtype=np.float64).reshape(-1)
            xe = np.zeros(F)
            pending.append((pipe.submit_fit(np.ascontiguousarray(positions, dtype=np.float64),
                                            np.ascontiguousarray(numbers, dtype=np.int32), images, y, xe),
                            xe, energy, len(positions)))
            if len(pending) == depth:
                finish(pending.pop(0))
        while pending:
            finish(pending.pop(0))
        gram_f, ord_f, moments = pipe.export_gram()
    finally:
        pipe.close()
    stats.gram_f += gram_f
    stats.ord_f += ord_f
    stats.moments[3:6] += moments
    return stats


def atom_range(n_atoms, rank, world_size):
    """Contiguous, balanced atom range (first, count) of `rank`."""
    base, rem = divmod(int(n_atoms), int(world_size))
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def all_reduce_partials(buffer, group=None):
    """Sum a rank's partial [forces (3N), energy] buffer over the group in place."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buffer, op=dist.ReduceOp.SUM, group=group)
    return buffer


class ShardedEvaluator:
    """Energy and forces of one frame with its atoms split over the ranks of `group`.

    Everything stays on the device: `positions` / `numbers` are CUDA tensors holding the
    WHOLE frame on every rank, the result is a CUDA tensor [3N + 1] (forces row-major, then
    the energy), identical on all ranks after the all-reduce."""

    def __init__(self, basis, coefficients, device=None, group=None):
        import torch
        import torch.distributed as dist
        from uf3_b200.engine import Engine
        self.group = group
        on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if on else 0
        self.world = dist.get_world_size(group) if on else 1
        # profiling aid: UF3B_FAKE_WORLD=k makes a single process do the work of rank 0 of k
        # (its share of the centres, no collective) so that one GPU shows the per-rank step
        self.fake_world = int(os.environ.get("UF3B_FAKE_WORLD", "0")) if self.world == 1 else 0
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.engine = Engine(basis, device=self.device, deferred_lists=True)
        self.engine.set_coefficients(coefficients)
        self.out = None

    def energy_forces(self, positions, numbers, images):
        import torch
        n = positions.shape[0]
        if self.out is None or self.out.numel() != 3 * n + 1:
            self.out = torch.empty(3 * n + 1, dtype=torch.float64, device=positions.device)
        stream = torch.cuda.current_stream().cuda_stream
        self.engine.build_neighbors_device(positions.data_ptr(), numbers.data_ptr(), n, images, stream,
                                           centres=atom_range(n, self.rank, self.fake_world or self.world))
        self.engine.energy_forces_device(self.out[3 * n:].data_ptr(), self.out.data_ptr(), stream)
        all_reduce_partials(self.out, self.group)
        return self.out[3 * n], self.out[:3 * n].view(n, 3)

    def close(self):
        self.engine.close()
