"""Helpers shared by the tests: load golden fixtures (tests/golden/*.npz, written by
oracle/make_golden.py from the running reference) and rebuild the product-side basis."""
import glob
import json
import os

import numpy as np

from uf3_b200 import bspline, composition, geometry
from uf3_b200.atoms import Atoms

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names(kind=None):
    names = sorted(os.path.splitext(os.path.basename(p))[0]
                   for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    names = [n for n in names if not n.startswith(("fit_", "rdf_"))]      # multi-frame fixtures have their own tests
    if kind == "featurize":     # dev_*: the documented deviation from the reference, tests/test_deviation_fixture.py
        names = [n for n in names if not n.startswith(("calc_", "dev_"))]
    elif kind == "calculator":
        names = [n for n in names if n.startswith("calc_")]
    return names


def _decode_keys(obj):
    if isinstance(obj, dict):
        out = {}
        for k, v in obj.items():
            if "-" in k:
                k = tuple(k.split("-"))
            elif k.isdigit():
                k = int(k)
            out[k] = _decode_keys(v)
        return out
    return obj


def _arrays(knots_map):
    out = {}
    for key, val in knots_map.items():
        if len(key) == 2:
            out[key] = np.array(val)
        else:
            out[key] = [np.array(v) for v in val]
    return out


class Case:
    def __init__(self, name):
        self.name = name
        data = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
        self.data = {k: data[k] for k in data.files}
        cfg = json.loads(str(self.data["config"]))
        kwargs = _decode_keys(cfg["kwargs"])
        if "knots_map" in kwargs:
            kwargs["knots_map"] = _arrays(kwargs["knots_map"])
        self.element_list = cfg["element_list"]
        self.degree = cfg["degree"]
        self.kwargs = kwargs
        self.positions = self.data["positions"]
        self.numbers = self.data["numbers"]
        self.cell = self.data["cell"]
        self.pbc = self.data["pbc"]

    def basis(self):
        chem = composition.ChemicalSystem(self.element_list, degree=self.degree)
        return bspline.BSplineBasis(chem, **self.kwargs)

    def atoms(self):
        return Atoms(numbers=self.numbers, positions=self.positions,
                     cell=self.cell, pbc=self.pbc)

    def image_offsets(self, basis):
        return geometry.image_table(self.cell, self.pbc, basis.r_cut)[1]

    def __getitem__(self, key):
        return self.data[key]

    def __contains__(self, key):
        return key in self.data


def csr_from_pairs(i, j, n):
    """(i, j) pair arrays (i ascending) -> CSR (offsets, j sorted per row)."""
    order = np.lexsort((j, i))
    i, j = i[order], j[order]
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.add.at(offsets, i + 1, 1)
    return np.cumsum(offsets), j


def rel_err(got, want):
    got, want = np.asarray(got, float), np.asarray(want, float)
    scale = np.max(np.abs(want)) if want.size else 0.0
    if scale == 0.0:
        return float(np.max(np.abs(got))) if got.size else 0.0
    return float(np.max(np.abs(got - want)) / scale)
