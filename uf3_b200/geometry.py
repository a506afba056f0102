"""Periodic-image bookkeeping (host side).

UF3 never uses the minimum-image convention: it replicates the cell into an
explicit block of ghost images and numbers every ghost
(`/root/reference/uf3/data/geometry.py:14-149`).  "Neighbor indices" in this
code base therefore mean *supercell indices*  `image_rank * n_atoms + atom`,
with the real atoms as image rank 0.  This module produces that image list
(order and offsets) with the same numpy/scipy expressions the reference uses,
so ghost coordinates are bit-identical; the CUDA neighbor kernel consumes the
offsets table and never builds the full supercell.

  images per axis   n = ceil(r_cut / height_axis)            geometry.py:54-83
  per-axis order    [0, 1, -1, 2, -2, ...]                   geometry.py:131-138
  image rank order  meshgrid(a, b, c) 'xy' flattened: b slowest, a, c fastest
                                                             geometry.py:108-114
  ghost position    positions + dot([ia, ib, ic], cell)      geometry.py:146-147
"""
import warnings

import numpy as np
from scipy import linalg

from uf3_b200.atoms import Atoms


def get_supercell_factors(cell, r_cut: float = 10):
    cell = np.asarray(cell, dtype=float)
    if np.all(cell == 0):
        return [1, 1, 1]
    if np.any(np.linalg.norm(cell, 2, axis=1) == 0):
        warnings.warn("Unit cell has 0-length lattice vector(s).")
        return [1, 1, 1]
    a, b, c = cell
    normals = [np.cross(b, c), np.cross(a, c), np.cross(a, b)]
    heights = [n * np.dot(v, n) / np.dot(n, n) for v, n in zip((a, b, c), normals)]
    return np.ceil([r_cut / linalg.norm(h) for h in heights])


def generate_periodic_image_indices(cell, r_cut):
    per_axis = []
    for n in get_supercell_factors(cell, r_cut):
        ring = np.repeat(np.arange(n + 1), 2)[1:]
        ring[::2] *= -1
        per_axis.append(ring)
    return per_axis


def image_table(cell, pbc, r_cut):
    """(image_abc (n_img, 3) int64, offsets (n_img, 3) float64); row 0 is the home cell."""
    cell = np.asarray(cell, dtype=float).reshape(3, 3)
    if not np.any(pbc):
        return np.zeros((1, 3), dtype=np.int64), np.zeros((1, 3))
    axes = generate_periodic_image_indices(cell, r_cut)
    for dim in range(3):
        if not pbc[dim]:
            axes[dim] = axes[dim][:1]
    grids = np.meshgrid(*axes)
    abc = np.stack([g.flatten() for g in grids], axis=1)
    offsets = np.array([np.dot([ia, ib, ic], cell) for ia, ib, ic in abc],
                       dtype=np.float64).reshape(-1, 3)
    return abc.astype(np.int64), offsets


def get_supercell(geometry, r_cut: float = 10, sort_indices: bool = False):
    """Explicit ghost supercell as an `Atoms` (API parity; O(M) memory).

    The GPU path does not call this — it is for inspection and for callers that
    pass `supercell=` through the featurizer API."""
    if sort_indices:
        raise NotImplementedError("sort_indices=True is not used on the UF3 hot path")
    positions = np.asarray(geometry.get_positions())
    numbers = np.asarray(geometry.get_atomic_numbers())
    _, offsets = image_table(geometry.get_cell(), geometry.get_pbc(), r_cut)
    sup_positions = (positions[None, :, :] + offsets[:, None, :]).reshape(-1, 3)
    return Atoms(numbers=np.tile(numbers, len(offsets)), positions=sup_positions)
