"""Minimal atomic-configuration container and frame extraction (host side).

The reference's data contract is `ase.Atoms` in, numpy out.
ASE is optional here: `Atoms` below offers the subset of the ASE interface the
UF3 path uses (`get_positions`, `get_atomic_numbers`, `get_cell`, `get_pbc`,
`get_chemical_symbols`, `pbc`, `calc`), and `frame_arrays` accepts either this
class or a real `ase.Atoms` (duck-typed).
"""
import numpy as np

from uf3_b200 import elements


class Atoms:
    def __init__(self, symbols=None, positions=None, numbers=None, cell=None,
                 pbc=None, calculator=None, info=None):
        if numbers is None:
            numbers = elements.symbols2numbers(symbols) if symbols is not None else []
        self.numbers = np.array(numbers, dtype=np.int64)
        n = len(self.numbers)
        self.positions = (np.zeros((n, 3)) if positions is None
                          else np.array(positions, dtype=np.float64).reshape(n, 3))
        self.set_cell(cell)
        self.set_pbc(pbc)
        self.calc = calculator
        self.info = dict(info or {})

    def __len__(self):
        return len(self.numbers)

    # -- setters ---------------------------------------------------------
    def set_cell(self, cell, scale_atoms=False):
        new = np.zeros((3, 3)) if cell is None else np.array(cell, dtype=np.float64)
        if new.shape == (3,):
            new = np.diag(new)
        if new.shape != (3, 3):
            raise ValueError("cell must be None, 3 lengths or a 3x3 array")
        if scale_atoms and hasattr(self, "cell"):
            self.positions = self.positions @ np.linalg.solve(self.cell, new)
        self.cell = new

    def set_pbc(self, pbc):
        if pbc is None:
            pbc = False
        if np.ndim(pbc) == 0:
            pbc = [bool(pbc)] * 3
        self.pbc = np.array(pbc, dtype=bool)

    def set_positions(self, positions):
        self.positions = np.array(positions, dtype=np.float64).reshape(len(self), 3)

    def set_calculator(self, calc):
        self.calc = calc

    # -- getters ---------------------------------------------------------
    def get_positions(self):
        return self.positions.copy()

    def get_atomic_numbers(self):
        return self.numbers.copy()

    def get_chemical_symbols(self):
        return [elements.chemical_symbols[z] for z in self.numbers]

    def get_cell(self):
        return self.cell.copy()

    def get_pbc(self):
        return self.pbc.copy()

    def get_volume(self):
        return float(abs(np.linalg.det(self.cell)))

    def copy(self):
        return Atoms(numbers=self.numbers, positions=self.positions,
                     cell=self.cell, pbc=self.pbc, info=self.info)

    # -- calculator protocol (what ASE's Atoms forwards to .calc) ---------
    def get_potential_energy(self, force_consistent=False):
        return self.calc.get_potential_energy(self)

    def get_forces(self):
        return self.calc.get_forces(self)

    def get_stress(self):
        return self.calc.get_stress(self)


def frame_arrays(geom):
    """(positions (N,3) f64 C-contig, numbers (N,) i32, cell (3,3) f64, pbc (3,) u8)."""
    positions = np.ascontiguousarray(geom.get_positions(), dtype=np.float64)
    numbers = np.ascontiguousarray(geom.get_atomic_numbers(), dtype=np.int32)
    cell = np.ascontiguousarray(np.array(geom.get_cell()), dtype=np.float64).reshape(3, 3)
    pbc = np.ascontiguousarray(np.array(geom.get_pbc(), dtype=bool).astype(np.uint8))
    if positions.ndim != 2 or positions.shape[1] != 3:
        raise ValueError("positions must have shape (n_atoms, 3)")
    if len(numbers) != len(positions):
        raise ValueError("numbers and positions disagree on n_atoms")
    return positions, numbers, cell, pbc
