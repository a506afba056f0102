import numpy as np


class Cell(np.ndarray):
    """Stand-in for ase.cell.Cell: a (3,3) float ndarray subclass."""
    def __new__(cls, array):
        return np.asarray(array, dtype=float).reshape(3, 3).view(cls)
