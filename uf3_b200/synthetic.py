"""Seeded synthetic frames and the basis settings of BASELINE.json's configs.

Positions = ideal lattice + N(0, sigma^2) per coordinate (not wrapped), orthorhombic cell,
fully periodic; `numpy.random.default_rng(seed)`.
"""
import numpy as np

from uf3_b200 import bspline, composition

# examples/tungsten_extxyz/uf23_potential_demo.ipynb:319-327 (F = 73)
W_DEMO = dict(r_min_map={("W", "W"): 0.001, ("W", "W", "W"): [1.5, 1.5, 1.5]},
              r_max_map={("W", "W"): 5.5, ("W", "W", "W"): [3.5, 3.5, 7.0]},
              resolution_map={("W", "W"): 15, ("W", "W", "W"): [6, 6, 12]},
              leading_trim={2: 0, 3: 3}, trailing_trim={2: 3, 3: 3})
# uf23_potential_demo.ipynb:281 / supplement/potentials/manuscript_uf23.json (F = 456)
W_MANUSCRIPT = dict(r_min_map={("W", "W"): 1.5, ("W", "W", "W"): [1.5, 1.5, 1.5]},
                    r_max_map={("W", "W"): 5.5, ("W", "W", "W"): [4.25, 4.25, 8.5]},
                    resolution_map={("W", "W"): 22, ("W", "W", "W"): [10, 10, 20]},
                    leading_trim={2: 0, 3: 3}, trailing_trim={2: 3, 3: 3})
# examples/NeXe_lammps/model_pair.json (2-body, 3 pair types, F = 56)
NEXE_PAIR = dict(r_min_map={("Ne", "Ne"): 2.0, ("Ne", "Xe"): 2.5, ("Xe", "Xe"): 3.0},
                 r_max_map={("Ne", "Ne"): 6.0, ("Ne", "Xe"): 7.0, ("Xe", "Xe"): 8.0},
                 resolution_map={("Ne", "Ne"): 15, ("Ne", "Xe"): 15, ("Xe", "Xe"): 15})


def w_basis(kind="demo"):
    chem = composition.ChemicalSystem(["W"], degree=3)
    return bspline.BSplineBasis(chem, **(W_DEMO if kind == "demo" else W_MANUSCRIPT))


def nexe_basis():
    chem = composition.ChemicalSystem(["Ne", "Xe"], degree=2)
    return bspline.BSplineBasis(chem, **NEXE_PAIR)


def _cells(reps):
    grid = np.indices(reps).reshape(3, -1).T
    return grid.astype(np.float64)


def bcc_w(reps=(10, 20, 25), a=3.165, sigma=0.05, seed=0):
    """(positions (N,3), numbers (N,), cell (3,3), pbc (3,)); N = 2 * prod(reps)."""
    rng = np.random.default_rng(seed)
    base = np.array([[0, 0, 0], [0.5, 0.5, 0.5]])
    pos = (_cells(reps)[:, None, :] + base[None, :, :]).reshape(-1, 3) * a
    pos = pos + rng.normal(0, sigma, pos.shape)
    return pos, np.full(len(pos), 74, dtype=np.int32), np.diag(np.array(reps) * a), np.ones(3, bool)


def nexe(reps=(25, 25, 10), a=8.0, sigma=0.1, seed=0):
    """fcc Ne + fcc Xe shifted by a/2 along x (examples/NeXe_lammps/dataset/generate_Ne-Xe.in:6-12);
    N = 8 * prod(reps)."""
    rng = np.random.default_rng(seed)
    fcc = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    ne = (_cells(reps)[:, None, :] + fcc[None]).reshape(-1, 3)
    xe = ne + np.array([0.5, 0, 0])
    pos = np.concatenate([ne, xe]) * a
    pos = pos + rng.normal(0, sigma, pos.shape)
    numbers = np.array([10] * len(ne) + [54] * len(xe), dtype=np.int32)
    return pos, numbers, np.diag(np.array(reps) * a), np.ones(3, bool)


# tests/test_representation.py:605-648 of the reference (the Fe-C basis of its committed golden rows,
# rattled_steel_features.json): six trios, two of them of symmetry 1; F = 609
_FEC_PAIRS = [("Fe", "Fe"), ("Fe", "C"), ("C", "C")]
_FEC_TRIOS = [("Fe", "Fe", "Fe"), ("Fe", "Fe", "C"), ("Fe", "C", "C"),
              ("C", "Fe", "Fe"), ("C", "Fe", "C"), ("C", "C", "C")]
FEC = dict(r_min_map={**{p: 0.1 for p in _FEC_PAIRS}, **{t: [1.5, 1.5, 1.5] for t in _FEC_TRIOS}},
           r_max_map={**{p: 6.0 for p in _FEC_PAIRS}, **{t: [5.0, 5.0, 10.0] for t in _FEC_TRIOS}},
           resolution_map={**{p: 12 for p in _FEC_PAIRS}, **{t: [4, 4, 8] for t in _FEC_TRIOS}},
           knot_strategy="linear", offset_1b=True, leading_trim=0, trailing_trim=3)


def fec_basis():
    chem = composition.ChemicalSystem(["Fe", "C"], degree=3)
    return bspline.BSplineBasis(chem, **FEC)


def b2_fec(reps=(10, 20, 25), a=2.87, sigma=0.05, seed=0):
    """B2 (CsCl-type) Fe-C: Fe on the cube corners, C on the body centres; N = 2 * prod(reps).
    At a = 2.87 A every atom has 58 neighbours inside the 5 A three-body cutoff of `FEC`
    (1 653 triangles per centre, against 91 in bulk W with the demo basis)."""
    rng = np.random.default_rng(seed)
    cells = _cells(reps)
    pos = np.concatenate([cells, cells + 0.5]) * a
    pos = pos + rng.normal(0, sigma, pos.shape)
    numbers = np.array([26] * len(cells) + [6] * len(cells), dtype=np.int32)
    return pos, numbers, np.diag(np.array(reps) * a), np.ones(3, bool)
