"""Stand-in for the parts of `ase` the UF3 hot path touches (TEST INFRASTRUCTURE ONLY)."""
import copy as _copy
import numpy as np
from ase import symbols as _sym
from ase.cell import Cell

__version__ = "0.0-shim"


class Atoms:
    def __init__(self, symbols=None, positions=None, numbers=None, cell=None,
                 pbc=None, calculator=None, info=None):
        if symbols is not None and numbers is None:
            numbers = _sym.symbols2numbers(symbols)
        self.numbers = np.array(numbers if numbers is not None else [], dtype=int)
        n = len(self.numbers)
        if positions is None:
            positions = np.zeros((n, 3))
        self.positions = np.array(positions, dtype=float).reshape(n, 3)
        self.set_cell(cell)
        self.set_pbc(pbc)
        self.calc = calculator
        self.info = dict(info or {})
        self.arrays = {}

    def __len__(self):
        return len(self.numbers)

    def set_cell(self, cell, scale_atoms=False):
        if cell is None:
            cell = np.zeros((3, 3))
        cell = np.array(cell, dtype=float)
        if cell.shape == (3,):
            cell = np.diag(cell)
        old = getattr(self, "cell", None)
        self.cell = Cell(cell)
        if scale_atoms and old is not None:
            m = np.linalg.solve(np.asarray(old), np.asarray(self.cell))
            self.positions = self.positions @ m

    def set_pbc(self, pbc):
        if pbc is None:
            pbc = False
        if isinstance(pbc, (bool, np.bool_, int)):
            pbc = [bool(pbc)] * 3
        self.pbc = np.array(pbc, dtype=bool)

    def get_pbc(self):
        return self.pbc.copy()

    def get_cell(self):
        return Cell(np.array(self.cell))

    def get_positions(self):
        return self.positions.copy()

    def set_positions(self, positions):
        self.positions = np.array(positions, dtype=float).reshape(len(self), 3)

    def get_atomic_numbers(self):
        return self.numbers.copy()

    def get_chemical_symbols(self):
        return [_sym.chemical_symbols[z] for z in self.numbers]

    def get_volume(self):
        return abs(np.linalg.det(np.asarray(self.cell)))

    def copy(self):
        new = Atoms(numbers=self.numbers.copy(), positions=self.positions.copy(),
                    cell=np.array(self.cell), pbc=self.pbc.copy(),
                    info=_copy.deepcopy(self.info))
        return new

    def __delitem__(self, idx):
        mask = np.ones(len(self), dtype=bool)
        mask[np.asarray(idx, dtype=int)] = False
        self.numbers = self.numbers[mask]
        self.positions = self.positions[mask]

    def set_calculator(self, calc):
        self.calc = calc

    def get_potential_energy(self, force_consistent=False):
        return self.calc.get_potential_energy(self)

    def get_forces(self):
        return self.calc.get_forces(self)

    def get_stress(self):
        return self.calc.get_stress(self)
