"""Element tables and symbol handling (host side).

Stands in for the `ase.symbols` helpers the reference calls
(`/root/reference/uf3/data/composition.py:9-10,54,159`) so the package works
whether or not ASE is installed.
"""
import re
from typing import Iterable, List, Union

chemical_symbols = (
    "X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co "
    "Ni Cu Zn Ga Ge As Se Br Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te "
    "I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir "
    "Pt Au Hg Tl Pb Bi Po At Rn Fr Ra Ac Th Pa U Np Pu Am Cm Bk Cf Es Fm Md No "
    "Lr Rf Db Sg Bh Hs Mt Ds Rg Cn Nh Fl Mc Lv Ts Og").split()
atomic_numbers = {sym: z for z, sym in enumerate(chemical_symbols)}

_FORMULA_TOKEN = re.compile(r"([A-Z][a-z]?)(\d*)")


def number_of(element: Union[str, int]) -> int:
    """Atomic number of a symbol (or pass an int through)."""
    if isinstance(element, str):
        try:
            return atomic_numbers[element]
        except KeyError:
            raise ValueError(f"Unknown element symbol: {element!r}") from None
    return int(element)


def symbol_of(element: Union[str, int]) -> str:
    if isinstance(element, str):
        number_of(element)
        return element
    return chemical_symbols[int(element)]


def symbols2numbers(symbols: Union[str, Iterable]) -> List[int]:
    """'Fe8C3' / ['Fe', 'C'] / [26, 6] -> list of atomic numbers."""
    if isinstance(symbols, str):
        numbers = []
        for sym, count in _FORMULA_TOKEN.findall(symbols):
            numbers.extend([number_of(sym)] * (int(count) if count else 1))
        return numbers
    return [number_of(s) for s in symbols]
