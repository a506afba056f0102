"""Ridge and curvature penalty matrices for the regularised least-squares fit.

Host-side, tiny, built once per fit.  Same matrices as
`/root/reference/uf3/regression/regularize.py` (:16 ridge, :33 1-D curvature,
:61 block-diagonal combine, :150 3-D curvature), assembled with array
operations instead of per-entry loops.
"""
from typing import List

import numpy as np

DEFAULT_REGULARIZER_GRID = dict(ridge_1b=1e-16, ridge_2b=0.0, ridge_3b=1e-10,
                                curve_2b=1e-16, curve_3b=1e-16)


def get_ridge_penalty_matrix(n_features: int) -> np.ndarray:
    return np.eye(n_features)


def get_curvature_penalty_matrix_1D(n_features: int) -> np.ndarray:
    """Second-difference stencil (1, -2, 1); the two corner entries are -1."""
    n = n_features
    matrix = -2.0 * np.eye(n) + np.eye(n, k=1) + np.eye(n, k=-1)
    matrix[0, 0] = -1.0
    matrix[n - 1, n - 1] = -1.0
    return matrix


def _laplacian_stencil(shape, flatten):
    """Row r (C-order multi-index) holds +1 on each in-bounds axis neighbour
    and minus their count on the centre."""
    size = int(np.prod(shape))
    out = np.zeros((size,) + tuple(shape))
    centre = np.indices(shape).reshape(len(shape), size)
    rows = np.arange(size)
    degree = np.zeros(size)
    for axis, extent in enumerate(shape):
        for step in (-1, 1):
            moved = centre.copy()
            moved[axis] += step
            ok = (moved[axis] >= 0) & (moved[axis] < extent)
            out[(rows[ok],) + tuple(moved[:, ok])] = 1.0
            degree += ok
    out[(rows,) + tuple(centre)] = -degree
    return out.reshape(size, size) if flatten else out


def get_curvature_penalty_matrix_2D(L: int, M: int, flatten: bool = True):
    return _laplacian_stencil((L, M), flatten)


def get_curvature_penalty_matrix_3D(L: int, M: int, N: int, flatten: bool = True):
    return _laplacian_stencil((L, M, N), flatten)


def combine_regularizer_matrices(matrices: List) -> np.ndarray:
    """Block-diagonal stack; blocks need not be square (rows = conditions)."""
    n_rows = sum(m.shape[0] for m in matrices)
    n_cols = sum(m.shape[1] for m in matrices)
    full = np.zeros((n_rows, n_cols))
    r = c = 0
    for m in matrices:
        full[r:r + m.shape[0], c:c + m.shape[1]] = m
        r += m.shape[0]
        c += m.shape[1]
    return full
