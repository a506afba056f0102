"""Owner of the device-side handles (basis tables + one reusable neighbour list).

`Engine` is the only place in the package that calls the C ABI (`include/uf3b.h`).
`process.BasisFeaturizer` and `calculator.UFCalculator` hold one lazily, drop it when
pickled (the reference ships featurizers to worker processes, util/parallel.py:182) and
re-create it in the process / on the GPU where they are next used.
"""
import ctypes as C

import numpy as np

from uf3_b200 import _native, geometry
from uf3_b200.tables import BasisTables


def _ptr(arr):
    return C.c_void_p(arr.ctypes.data)


class Engine:
    def __init__(self, basis, device=None, frames_in_flight=1, deferred_lists=False):
        """`frames_in_flight` = k: this engine is one of k working on consecutive frames on k
        streams; its feature kernels then take 1/k of the SM resources per launch.
        `deferred_lists` (MD loops): a list build that can reuse the previous cell grid returns
        without a host wait; `energy_forces*` verifies it after queueing its own kernels and
        repeats build + evaluation in the rare case it was invalid (`uf3b_basis_set_deferred_lists`)."""
        self._lib = _native.lib()
        # the basis handle remembers the device it was created on and every C-ABI call that takes
        # it makes that device current for its duration (DeviceGuard), so engines on different
        # GPUs can be driven from one thread
        self.device = None if device is None else int(device)
        if device is not None:
            _native.check(self._lib.uf3b_set_device(int(device)))
        self.tables = BasisTables(basis)
        self.n_feats = self.tables.n_feats
        self._basis = C.c_void_p()
        _native.check(self._lib.uf3b_basis_create(C.byref(self.tables.desc), C.byref(self._basis)))
        if frames_in_flight != 1:
            _native.check(self._lib.uf3b_basis_set_frames_in_flight(self._basis, int(frames_in_flight)))
        self.deferred_lists = bool(deferred_lists)
        if self.deferred_lists:
            _native.check(self._lib.uf3b_basis_set_deferred_lists(self._basis, 1))
        self._last_build = None
        self._nlist = C.c_void_p()
        self.n_atoms = 0
        self.has_coefficients = False

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_lib", None) is None:
            return
        if self._nlist:
            self._lib.uf3b_nlist_destroy(self._nlist)
            self._nlist = C.c_void_p()
        if self._basis:
            self._lib.uf3b_basis_destroy(self._basis)
            self._basis = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ inputs
    def set_coefficients(self, coefficients):
        c = np.ascontiguousarray(coefficients, dtype=np.float64)
        _native.check(self._lib.uf3b_basis_set_coefficients(
            self._basis, c.ctypes.data_as(C.POINTER(C.c_double)), len(c)))
        self.has_coefficients = True

    def build_neighbors(self, positions, numbers, cell=None, pbc=None, images=None, stream=None,
                        centres=None):
        """Kernel A.  `images` = (abc (n_img,3) int, offsets (n_img,3) float64) overrides the
        periodic-image table derived from (cell, pbc, r_cut) by `geometry.image_table`.
        `centres` = (first, count) builds the rows of that atom range only (one frame split
        over ranks); `energy_forces` then returns this rank's partial sums."""
        positions = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
        numbers = np.ascontiguousarray(numbers, dtype=np.int32)
        if len(numbers) != len(positions):
            raise ValueError("numbers and positions disagree on n_atoms")
        if images is None:
            if cell is None or pbc is None or not np.any(pbc):
                images = (np.zeros((1, 3), dtype=np.int64), np.zeros((1, 3)))
            else:
                images = geometry.image_table(cell, pbc, self.tables.r_cut)
        abc = np.ascontiguousarray(images[0], dtype=np.int32).reshape(-1, 3)
        offsets = np.ascontiguousarray(images[1], dtype=np.float64).reshape(-1, 3)
        first, count = (0, len(positions)) if centres is None else centres
        _native.check(self._lib.uf3b_neighbors_build_range(
            self._basis, len(positions), _ptr(positions), _ptr(numbers), len(offsets),
            _ptr(offsets), _ptr(abc), int(first), int(count), C.byref(self._nlist), stream))
        self.n_atoms = len(positions)
        return self

    def build_neighbors_device(self, positions_ptr, numbers_ptr, n_atoms, images, stream=None, centres=None):
        """Same with DEVICE pointers for positions (n,3 float64) and numbers (n int32)."""
        cached = getattr(self, "_images_cache", None)
        if cached is not None and cached[0] is images:      # MD loops pass the same table every step
            abc, offsets = cached[1], cached[2]
        else:
            abc = np.ascontiguousarray(images[0], dtype=np.int32).reshape(-1, 3)
            offsets = np.ascontiguousarray(images[1], dtype=np.float64).reshape(-1, 3)
            self._images_cache = (images, abc, offsets)
        first, count = (0, int(n_atoms)) if centres is None else centres
        self._last_build = (positions_ptr, numbers_ptr, int(n_atoms), offsets, abc, int(first), int(count), stream)
        self._build_device()
        self.n_atoms = int(n_atoms)
        return self

    def _build_device(self):
        positions_ptr, numbers_ptr, n_atoms, offsets, abc, first, count, stream = self._last_build
        _native.check(self._lib.uf3b_neighbors_build_range(
            self._basis, n_atoms, C.c_void_p(positions_ptr), C.c_void_p(numbers_ptr),
            len(offsets), _ptr(offsets), _ptr(abc), first, count, C.byref(self._nlist), stream))

    def neighbor_count(self, which):
        """Number of entries in list 2 or 3 of the current configuration."""
        total = C.c_int64()
        _native.check(self._lib.uf3b_neighbors_count(self._nlist, which, C.byref(total)))
        return int(total.value)

    def neighbor_list(self, which):
        """CSR (offsets int64 [n+1], supercell indices int64) of list 2 or 3 (parity hook)."""
        total = C.c_int64()
        _native.check(self._lib.uf3b_neighbors_count(self._nlist, which, C.byref(total)))
        offsets = np.zeros(self.n_atoms + 1, dtype=np.int64)
        idx = np.zeros(max(total.value, 1), dtype=np.int64)
        _native.check(self._lib.uf3b_neighbors_export(
            self._nlist, which, offsets.ctypes.data_as(C.POINTER(C.c_int64)),
            idx.ctypes.data_as(C.POINTER(C.c_int64))))
        return offsets, idx[:total.value]

    # ------------------------------------------------------------------ kernels B
    def featurize(self, energy=True, forces=True, out_energy=None, out_forces=None, stream=None):
        """Feature rows of the configuration last passed to `build_neighbors`.

        Returns (x_energy [F] or None, x_forces [3N, F] or None); force row c*N + a.
        `out_*` may be preallocated (e.g. pinned) float64 arrays."""
        n, F = self.n_atoms, self.n_feats
        xe = xf = None
        if energy:
            xe = out_energy if out_energy is not None else np.empty(F)
        if forces:
            xf = out_forces if out_forces is not None else np.empty((3 * n, F))
            if xf.shape != (3 * n, F) or not xf.flags.c_contiguous or xf.dtype != np.float64:
                raise ValueError("out_forces must be a C-contiguous float64 (3N, F) array")
        _native.check(self._lib.uf3b_featurize(
            self._basis, self._nlist, _ptr(xe) if energy else None,
            _ptr(xf) if (forces and n) else None, F, stream))
        return xe, xf

    def featurize_device(self, energy_ptr, forces_ptr, ld, stream=None):
        """Device-pointer form: rows are left in device memory (row stride `ld` doubles)."""
        _native.check(self._lib.uf3b_featurize(
            self._basis, self._nlist, C.c_void_p(energy_ptr) if energy_ptr else None,
            C.c_void_p(forces_ptr) if forces_ptr else None, int(ld), stream))

    def _featurize_mixed(self, out_energy, forces_ptr, ld, stream=None):
        """Energy row into a HOST array, force rows (if any) left at a DEVICE address."""
        _native.check(self._lib.uf3b_featurize(
            self._basis, self._nlist, _ptr(out_energy),
            C.c_void_p(forces_ptr) if forces_ptr else None, int(ld), stream))

    def energy_forces(self, energy=True, forces=True, virial=False, stream=None):
        """(energy, forces [N,3]) of the current configuration; with `virial=True` also the
        3x3 strain derivative W = dE/d(strain) (stress = W / volume) as a third item."""
        n = self.n_atoms
        e = np.zeros(1) if energy else None
        f = np.zeros((n, 3)) if forces else None
        w = np.zeros((3, 3)) if virial else None
        _native.check(self._lib.uf3b_energy_forces(
            self._basis, self._nlist, _ptr(e) if energy else None,
            _ptr(f) if (forces and n) else None, _ptr(w) if virial else None, stream))
        if virial:
            return (float(e[0]) if energy else None), f, w
        return (float(e[0]) if energy else None), f

    def energy_forces_device(self, energy_ptr, forces_ptr, stream=None):
        """Device-pointer form (energy: 1 double, forces: [N,3]); asynchronous on `stream`."""
        for attempt in range(3):
            rc = self._lib.uf3b_energy_forces(
                self._basis, self._nlist, C.c_void_p(energy_ptr) if energy_ptr else None,
                C.c_void_p(forces_ptr) if forces_ptr else None, None, stream)
            if rc != _native.RETRY or self._last_build is None:
                break
            # the deferred list build was invalid (atoms left the cached grid / arrays regrown): the
            # library has dropped the grid or grown the arrays, so the next build is a checked one
            self._build_device()
        _native.check(rc)

    # ------------------------------------------------------------------ instrumentation
    def launch_count(self):
        return int(self._lib.uf3b_launch_count())

    def set_timing(self, enabled):
        self._lib.uf3b_set_timing(1 if enabled else 0)

    def last_kernel_ms(self):
        return float(self._lib.uf3b_last_kernel_ms())

    def probe_fp64_tflops(self):
        out = C.c_double()
        _native.check(self._lib.uf3b_probe_fp64_tflops(C.byref(out)))
        return float(out.value)
