"""Chemical system bookkeeping: element order, interaction order, hashes.

API mirror of `/root/reference/uf3/data/composition.py` (`ChemicalSystem` :28,
`sort_interaction_symbols` :191, `get_szudzik_hash` :293).  The reference orders
everything by a per-element key (`reference_X`, :12-25) that is the atomic
number for every element it lists; atomic number is used here directly.

Ordering contract (pins the feature-column layout, `bspline.py:525-575`):
  elements  ascending Z
  pairs     combinations-with-replacement of elements, each sorted by Z,
            listed lexicographically by (Z_a, Z_b)
  trios     (centre, j, k) with Z_j <= Z_k, lexicographic by (Z_c, Z_j, Z_k)
"""
import itertools
from typing import Any, Collection, Dict, List, Tuple

import numpy as np

from uf3_b200 import elements


def _key(symbol):
    return elements.number_of(symbol)


def sort_elements(symbols):
    return sorted(symbols, key=_key)


def sort_interaction_symbols(symbols: Collection[str], fix_first=True) -> Tuple:
    """Sort a pair by Z; for 3+ bodies keep the centre first and sort the rest."""
    symbols = list(symbols)
    if len(symbols) >= 3 and fix_first:
        return tuple([symbols[0]] + sort_elements(symbols[1:]))
    return tuple(sort_elements(symbols))


def sort_interaction_map(imap: Dict[Tuple[str], Any]) -> Dict[Tuple[str], Any]:
    return {sort_interaction_symbols(k): v for k, v in imap.items()}


def szudzik_pair(x, y):
    """Szudzik's pairing function, elementwise on integer arrays."""
    x = np.asarray(x, dtype=np.int64)
    y = np.asarray(y, dtype=np.int64)
    return np.where(x > y, x * x + y, y * y + x + y)


def get_szudzik_hash(array) -> np.ndarray:
    """Fold the pairing function left-to-right over the columns of `array`."""
    array = np.atleast_2d(np.asarray(array, dtype=np.int64))
    acc = array[:, 0]
    for col in range(1, array.shape[1]):
        acc = szudzik_pair(acc, array[:, col])
    return acc


class ChemicalSystem:
    def __init__(self, element_list: Collection, degree: int = 2) -> None:
        self.degree = int(degree)
        symbols = {elements.symbol_of(el) for el in element_list}
        self.element_list = tuple(sort_elements(symbols))
        self.numbers = [elements.number_of(el) for el in self.element_list]
        self.interactions_map = self.get_interactions_map()
        self.interactions = self.get_interactions_list()
        self.interaction_hashes = self.get_interaction_hashes()

    @staticmethod
    def from_config(config):
        return ChemicalSystem.from_dict(config)

    @staticmethod
    def from_dict(config: Dict[Any, Any]):
        return ChemicalSystem(element_list=config["element_list"],
                              degree=config["degree"])

    def as_dict(self):
        return dict(element_list=self.element_list, degree=self.degree)

    def __repr__(self):
        lines = ["ChemicalSystem:",
                 f"    Elements: {self.element_list}",
                 f"    Degree: {self.degree}",
                 f"    Pairs: {self.interactions_map[2]}"]
        if self.degree > 2:
            lines.append(f"    Trios: {self.interactions_map[3]}")
        return "\n".join(lines)

    def get_composition_tuple(self, geometry) -> np.ndarray:
        numbers = np.asarray(geometry.get_atomic_numbers())
        return np.array([np.count_nonzero(numbers == z) for z in self.numbers],
                        dtype=int)

    def get_interactions_map(self) -> Dict[int, Collection[Tuple[str]]]:
        els = list(self.element_list)
        imap = {1: self.element_list}
        imap[2] = [tuple(p) for p
                   in itertools.combinations_with_replacement(els, 2)]
        for d in range(3, self.degree + 1):
            combos = []
            for centre in els:
                for rest in itertools.combinations_with_replacement(els, d - 1):
                    combos.append((centre,) + tuple(rest))
            imap[d] = combos
        return imap

    def get_interactions_list(self) -> List:
        out = list(self.element_list)
        for d in range(2, self.degree + 1):
            out.extend(self.interactions_map[d])
        return out

    def get_interaction_hashes(self) -> Dict[int, np.ndarray]:
        hashes = {}
        for d in range(2, self.degree + 1):
            numbers = np.array([[elements.number_of(s) for s in combo]
                                for combo in self.interactions_map[d]],
                               dtype=np.int64)
            numbers[:, 1:] = np.sort(numbers[:, 1:], axis=1)
            hashes[d] = get_szudzik_hash(numbers)
        return hashes
