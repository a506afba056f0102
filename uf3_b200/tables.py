"""Flatten a `BSplineBasis` into the plain arrays of `uf3b_basis_desc` (include/uf3b.h).

Everything the kernels need from the host-side basis object:
  - knot vectors and strict pair bounds per pair interaction (distances.py:60-66)
  - three knot vectors per trio interaction
  - feature-column offsets following bspline.py:525-575 (elements, pairs, trios), counted
    WITHOUT the leading "y" column
  - the 3-body compression (`compress_3B`, bspline.py:664-690) as a map
    full-grid bin -> (compressed column, weight): a unit value in bin b of the (L, M, N)
    grid lands in column k with weight  flat_weights[k] * #{permutations of the trio's
    symmetry group that carry the canonical cell template_mask[k] onto b}.
"""
import ctypes as C
import itertools

import numpy as np

from uf3_b200 import elements
from uf3_b200._native import BasisDesc


def _symmetry_group(order):
    if order == 2:
        return [(0, 1, 2), (1, 0, 2)]
    if order == 3:
        return list(itertools.permutations(range(3)))
    return [(0, 1, 2)]


def bin_map(basis, trio):
    """(bin_col int32 [L*M*N], bin_weight float64 [L*M*N]) for one trio interaction."""
    shape = tuple(len(k) - 4 for k in basis.knots_map[trio])
    size = int(np.prod(shape))
    mask = np.asarray(basis.template_mask[trio], dtype=np.int64)
    weights = np.asarray(basis.flat_weights[trio], dtype=np.float64)
    col = np.full(size, -1, dtype=np.int32)
    w = np.zeros(size, dtype=np.float64)
    flat = np.arange(size, dtype=np.int64).reshape(shape)
    for perm in _symmetry_group(basis.symmetry[trio]):
        # compress_3B adds grid.transpose(perm): cell c of the folded grid reads bin src[c]
        view = flat.transpose(perm)
        if view.shape != shape:
            raise ValueError(f"{trio}: symmetry {basis.symmetry[trio]} needs equal grid extents")
        src = view.reshape(-1)[mask]
        col[src] = np.arange(len(mask), dtype=np.int32)
        np.add.at(w, src, weights)
    return col, w


class BasisTables:
    """Plain-array form of a basis + the ctypes descriptor that points into it."""

    def __init__(self, basis):
        self.element_list = list(basis.element_list)
        ne = len(self.element_list)
        self.numbers = np.array([elements.number_of(e) for e in self.element_list], dtype=np.int32)
        if np.any(np.diff(self.numbers) <= 0):
            raise ValueError("element_list must be sorted by atomic number")
        pairs = list(basis.interactions_map[2])
        expect = [(self.element_list[a], self.element_list[b])
                  for a in range(ne) for b in range(a, ne)]
        if [tuple(p) for p in pairs] != expect:
            raise ValueError("pair interactions are not in triangular order")
        trios = list(basis.interactions_map.get(3, [])) if basis.degree > 2 else []
        if trios:
            expect3 = [(self.element_list[c],) + p for c in range(ne) for p in expect]
            if [tuple(t) for t in trios] != expect3:
                raise ValueError("trio interactions are not in (centre, pair) order")
        sizes, offsets = basis.get_interaction_partitions()
        self.pairs, self.trios = pairs, trios
        self.n_feats = int(basis.n_feats)
        self.r_cut = float(basis.r_cut)
        self.partition = {key: (int(offsets[key]), int(sizes[key])) for key in pairs + trios}

        knots2 = [np.ascontiguousarray(basis.knots_map[p], dtype=np.float64) for p in pairs]
        self.pair_n_knots = np.array([len(k) for k in knots2], dtype=np.int32)
        self.pair_knots = np.concatenate(knots2)
        self.pair_r_min = np.array([basis.r_min_map[p] for p in pairs], dtype=np.float64)
        self.pair_r_max = np.array([basis.r_max_map[p] for p in pairs], dtype=np.float64)
        self.pair_col = np.array([offsets[p] for p in pairs], dtype=np.int32)
        for p, k, size in zip(pairs, knots2, (sizes[p] for p in pairs)):
            if len(k) - 4 != size:
                raise ValueError(f"{p}: {len(k)} knots do not give {size} basis functions")
        self.trims = (int(basis.leading_trim[2]), int(basis.trailing_trim[2]),
                      int(basis.leading_trim.get(3, 0)), int(basis.trailing_trim.get(3, 0)))

        if trios:
            knots3 = [np.ascontiguousarray(k, dtype=np.float64)
                      for t in trios for k in basis.knots_map[t]]
            self.trio_n_knots = np.array([len(k) for k in knots3], dtype=np.int32)
            self.trio_knots = np.concatenate(knots3)
            self.trio_col = np.array([offsets[t] for t in trios], dtype=np.int32)
            self.trio_n_cols = np.array([sizes[t] for t in trios], dtype=np.int32)
            self.trio_symmetry = np.array([basis.symmetry[t] for t in trios], dtype=np.int32)
            maps = [bin_map(basis, t) for t in trios]
            self.bin_col = np.concatenate([m[0] for m in maps])
            self.bin_weight = np.concatenate([m[1] for m in maps])
            self.trio_grid_offset = np.concatenate(
                [[0], np.cumsum([len(m[0]) for m in maps])[:-1]]).astype(np.int64)
        else:
            self.trio_n_knots = np.zeros(1, dtype=np.int32)
            self.trio_knots = np.zeros(1)
            self.trio_col = self.trio_n_cols = self.trio_symmetry = np.zeros(1, dtype=np.int32)
            self.bin_col = np.zeros(1, dtype=np.int32)
            self.bin_weight = np.zeros(1)
            self.trio_grid_offset = np.zeros(1, dtype=np.int64)

        d = BasisDesc()
        d.n_elements = ne
        d.atomic_numbers = self._p(self.numbers, C.c_int32)
        d.n_feats = self.n_feats
        (d.leading_trim_2b, d.trailing_trim_2b,
         d.leading_trim_3b, d.trailing_trim_3b) = self.trims
        d.pair_n_knots = self._p(self.pair_n_knots, C.c_int32)
        d.pair_knots = self._p(self.pair_knots, C.c_double)
        d.pair_r_min = self._p(self.pair_r_min, C.c_double)
        d.pair_r_max = self._p(self.pair_r_max, C.c_double)
        d.pair_col = self._p(self.pair_col, C.c_int32)
        d.n_trios = len(trios)
        d.trio_n_knots = self._p(self.trio_n_knots, C.c_int32)
        d.trio_knots = self._p(self.trio_knots, C.c_double)
        d.trio_col = self._p(self.trio_col, C.c_int32)
        d.trio_n_cols = self._p(self.trio_n_cols, C.c_int32)
        d.bin_col = self._p(self.bin_col, C.c_int32)
        d.bin_weight = self._p(self.bin_weight, C.c_double)
        d.trio_symmetry = self._p(self.trio_symmetry, C.c_int32)
        self.desc = d

    @staticmethod
    def _p(arr, ctype):
        return arr.ctypes.data_as(C.POINTER(ctype))

    def decompressed_grid(self, coefficients, trio_index):
        """`decompress_3B` of a trio's coefficients through the bin map (for tests)."""
        start = int(self.trio_grid_offset[trio_index])
        end = (int(self.trio_grid_offset[trio_index + 1])
               if trio_index + 1 < len(self.trios) else len(self.bin_col))
        col = self.bin_col[start:end]
        c = np.asarray(coefficients, dtype=np.float64)
        base = int(self.trio_col[trio_index])
        return np.where(col >= 0, c[base + np.maximum(col, 0)] * self.bin_weight[start:end], 0.0)
