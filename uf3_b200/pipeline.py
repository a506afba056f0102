"""Frame pipeline: feature rows of a stream of configurations with the device→host copy of
frame k overlapped with the kernels of frame k+1.

The reference produces one frame at a time and hands numpy rows to pandas
(`process.py:121-174`).  On the GPU the rows of a 10 000-atom frame are 17.5 MB (F = 73) to
110 MB (F = 456); copying them out synchronously costs 25–60 % of the step.  `FramePipeline`
keeps `depth` slots, each with its own engine (neighbour-list handle), compute stream, row
buffer on the device and pinned buffer on the host: the kernels of a frame run on the slot's
stream through the device-pointer form of the C ABI (`uf3b_neighbors_build`,
`uf3b_featurize`), a copy stream moves the finished rows out, and while the host waits on
the two small read-backs inside one frame's list build the GPU runs the other slot's feature
kernel.  torch is used for streams, events and pinned / device buffers only.
"""
import numpy as np
import torch

from uf3_b200 import geometry


class FramePipeline:
    def __init__(self, basis, max_atoms, device=None, forces=True, depth=2, frames_in_flight=1):
        from uf3_b200.engine import Engine
        index = torch.cuda.current_device() if device is None else int(device)
        self.engines = [Engine(basis, device=index, frames_in_flight=frames_in_flight) for _ in range(depth)]
        self.F = self.engines[0].n_feats
        self.max_atoms = int(max_atoms)
        self.forces = forces
        self.dev = torch.device("cuda", index)
        self.compute = [torch.cuda.Stream(self.dev) for _ in range(depth)]
        self.copy = torch.cuda.Stream(self.dev)
        self.depth = depth
        rows = 3 * self.max_atoms if forces else 0
        self.d_xf = [torch.empty((rows, self.F), dtype=torch.float64, device=self.dev) for _ in range(depth)]
        self.d_xe = [torch.empty(self.F, dtype=torch.float64, device=self.dev) for _ in range(depth)]
        self.h_xf = [torch.empty((rows, self.F), dtype=torch.float64).pin_memory() for _ in range(depth)]
        self.h_xe = [torch.empty(self.F, dtype=torch.float64).pin_memory() for _ in range(depth)]
        self.d_pos = [torch.empty((self.max_atoms, 3), dtype=torch.float64, device=self.dev) for _ in range(depth)]
        self.d_num = [torch.empty(self.max_atoms, dtype=torch.int32, device=self.dev) for _ in range(depth)]
        self.computed = [torch.cuda.Event() for _ in range(depth)]
        self.copied = [torch.cuda.Event() for _ in range(depth)]
        self.n_atoms = [0] * depth
        self.busy = [False] * depth
        self.turn = 0

    def submit(self, positions, numbers, images):
        """Queue one frame (host arrays, ideally pinned); returns the slot to read it from.
        `images` = geometry.image_table(cell, pbc, r_cut)."""
        slot = self.turn
        self.turn = (self.turn + 1) % self.depth
        n = len(positions)
        if n > self.max_atoms:
            raise ValueError("frame larger than the pipeline's max_atoms")
        if self.busy[slot]:
            self.copied[slot].synchronize()          # rows of the previous user of this slot are out
        self.busy[slot] = True
        self.n_atoms[slot] = n
        compute, eng = self.compute[slot], self.engines[slot]
        with torch.cuda.stream(compute):
            self.d_pos[slot][:n].copy_(torch.from_numpy(positions), non_blocking=True)
            self.d_num[slot][:n].copy_(torch.from_numpy(numbers), non_blocking=True)
            stream = compute.cuda_stream
            eng.build_neighbors_device(self.d_pos[slot].data_ptr(), self.d_num[slot].data_ptr(), n,
                                       images, stream)
            eng.featurize_device(self.d_xe[slot].data_ptr(),
                                 self.d_xf[slot].data_ptr() if self.forces else None, self.F, stream)
            self.computed[slot].record(compute)
        with torch.cuda.stream(self.copy):
            self.copy.wait_event(self.computed[slot])
            if self.forces:
                self.h_xf[slot][:3 * n].copy_(self.d_xf[slot][:3 * n], non_blocking=True)
            self.h_xe[slot].copy_(self.d_xe[slot], non_blocking=True)
            self.copied[slot].record(self.copy)
        return slot

    def result(self, slot):
        """(x_energy [F], x_forces [3N, F]) views of the pinned buffers of `slot`; valid until
        the slot is reused `depth` submissions later."""
        self.copied[slot].synchronize()
        n = self.n_atoms[slot]
        xf = self.h_xf[slot][:3 * n].numpy() if self.forces else None
        return self.h_xe[slot].numpy(), xf

    def drain(self):
        for slot in range(self.depth):
            if self.busy[slot]:
                self.copied[slot].synchronize()

    def launch_count(self):
        return self.engines[0].launch_count()

    def close(self):
        self.drain()
        for eng in self.engines:
            eng.close()


class NativePipeline:
    """`depth` frames in flight through HOST buffers, driven by the worker threads of the C
    library (`uf3b_pipeline_*`): submit() hands over pinned input / output arrays and returns a
    ticket at once, wait(ticket) blocks until that frame's rows are in its output arrays.
    No torch on this path; the caller owns every buffer and keeps it alive until wait()."""

    def __init__(self, basis, depth=3, device=None):
        import ctypes as C
        from uf3_b200 import _native
        from uf3_b200.tables import BasisTables
        self._C, self._native = C, _native
        self._lib = _native.lib()
        if device is not None:
            _native.check(self._lib.uf3b_set_device(int(device)))
        self.tables = BasisTables(basis)
        self.n_feats = self.tables.n_feats
        self.depth = int(depth)
        self._pipe = C.c_void_p()
        _native.check(self._lib.uf3b_pipeline_create(C.byref(self.tables.desc), self.depth, C.byref(self._pipe)))
        self._keep = {}

    def submit(self, positions, numbers, images, out_energy, out_forces):
        """positions (n,3) float64, numbers (n,) int32, images = geometry.image_table(...),
        out_energy (F,) and out_forces (3n, F) float64 C-contiguous (or None)."""
        C = self._C
        abc = np.ascontiguousarray(images[0], dtype=np.int32).reshape(-1, 3)
        offsets = np.ascontiguousarray(images[1], dtype=np.float64).reshape(-1, 3)
        if positions.dtype != np.float64 or numbers.dtype != np.int32 or not positions.flags.c_contiguous:
            raise ValueError("positions must be C-contiguous float64 and numbers int32")
        ticket = C.c_int64()
        self._native.check(self._lib.uf3b_pipeline_submit(
            self._pipe, len(positions), C.c_void_p(positions.ctypes.data), C.c_void_p(numbers.ctypes.data),
            len(offsets), C.c_void_p(offsets.ctypes.data), C.c_void_p(abc.ctypes.data),
            C.c_void_p(out_energy.ctypes.data) if out_energy is not None else None,
            C.c_void_p(out_forces.ctypes.data) if out_forces is not None else None,
            self.n_feats, C.byref(ticket)))
        self._keep[ticket.value % self.depth] = (positions, numbers, out_energy, out_forces)
        return ticket.value

    def wait(self, ticket):
        self._native.check(self._lib.uf3b_pipeline_wait(self._pipe, int(ticket)))

    def submit_fit(self, positions, numbers, images, y_forces, out_energy):
        """Fit job: the frame's force rows stay in HBM and go straight into the slot's normal-equation
        accumulator (`uf3b_pipeline_submit_fit`).  y_forces: (3n,) float64 targets in row order
        (fx_0.., fy_0.., fz_0..) or None; out_energy: (F,) float64 array receiving the energy row."""
        C = self._C
        abc = np.ascontiguousarray(images[0], dtype=np.int32).reshape(-1, 3)
        offsets = np.ascontiguousarray(images[1], dtype=np.float64).reshape(-1, 3)
        if positions.dtype != np.float64 or numbers.dtype != np.int32 or not positions.flags.c_contiguous:
            raise ValueError("positions must be C-contiguous float64 and numbers int32")
        if y_forces is not None and (y_forces.dtype != np.float64 or y_forces.size != 3 * len(positions)
                                     or not y_forces.flags.c_contiguous):
            raise ValueError("y_forces must be C-contiguous float64 with one target per force row")
        ticket = C.c_int64()
        self._native.check(self._lib.uf3b_pipeline_submit_fit(
            self._pipe, len(positions), C.c_void_p(positions.ctypes.data), C.c_void_p(numbers.ctypes.data),
            len(offsets), C.c_void_p(offsets.ctypes.data), C.c_void_p(abc.ctypes.data),
            C.c_void_p(y_forces.ctypes.data) if y_forces is not None else None,
            C.c_void_p(out_energy.ctypes.data) if out_energy is not None else None, C.byref(ticket)))
        self._keep[ticket.value % self.depth] = (positions, numbers, y_forces, out_energy)
        return ticket.value

    def export_gram(self):
        """(gram_f [F, F], ord_f [F], (n, sum y, sum y^2)) summed over the slots; waits for every frame."""
        C = self._C
        F = self.n_feats
        gram, ordinate, moments = np.zeros((F, F)), np.zeros(F), np.zeros(3)
        self._native.check(self._lib.uf3b_pipeline_export_gram(
            self._pipe, C.c_void_p(gram.ctypes.data), C.c_void_p(ordinate.ctypes.data),
            C.c_void_p(moments.ctypes.data)))
        return gram, ordinate, moments

    def close(self):
        if getattr(self, "_pipe", None):
            self._lib.uf3b_pipeline_destroy(self._pipe)
            self._pipe = self._C.c_void_p()
            self._keep = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def featurize_frames(featurizer, frames, max_atoms=None):
    """Generator of (x_energy, x_forces) copies for an iterable of geometries, pipelined."""
    from uf3_b200.atoms import frame_arrays
    frames = list(frames)
    if not frames:
        return
    if max_atoms is None:
        max_atoms = max(len(g) for g in frames)
    pipe = FramePipeline(featurizer.bspline_config, max_atoms, device=featurizer.device,
                         forces=featurizer.fit_forces)
    pending = []
    for geom in frames:
        positions, numbers, cell, pbc = frame_arrays(geom)
        images = geometry.image_table(cell, pbc, featurizer.r_cut) if np.any(pbc) else \
            (np.zeros((1, 3), dtype=np.int64), np.zeros((1, 3)))
        pending.append(pipe.submit(positions, numbers, images))
        if len(pending) == pipe.depth:
            xe, xf = pipe.result(pending.pop(0))
            yield xe.copy(), (xf.copy() if xf is not None else None)
    for slot in pending:
        xe, xf = pipe.result(slot)
        yield xe.copy(), (xf.copy() if xf is not None else None)
    pipe.close()
