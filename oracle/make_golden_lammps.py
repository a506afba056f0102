#!/usr/bin/env python
"""Golden LAMMPS hand-off texts written by the UNMODIFIED reference (build container only):

    python oracle/make_golden_lammps.py     # writes tests/golden/lammps_*.txt

`write_uf3_lammps_pot_files` is taken from the reference's script
(lammps_plugin/scripts/generate_uf3_lammps_pots.py:58-165) by compiling that function's own
source (the script's top-level imports need pymatgen, which is not installed), and
`export_tabulated_potential` from uf3/forcefield/lammps.py:218-271 the same way.

TEST INFRASTRUCTURE ONLY.
"""
import ast
import os
import sys
import tempfile
import warnings
from datetime import datetime

import numpy as np
from scipy import interpolate

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(HERE, "ref_shims"), "/root/reference"]
warnings.simplefilter("ignore")

import ase  # noqa: E402 (stand-in)
from uf3.regression import least_squares  # noqa: E402


def function_from(path, name, namespace):
    tree = ast.parse(open(path).read())
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    module = ast.Module(body=[node], type_ignores=[])
    exec(compile(module, path, "exec"), namespace)
    return namespace[name]


def main():
    out = os.path.join(REPO, "tests", "golden")
    writer = function_from("/root/reference/lammps_plugin/scripts/generate_uf3_lammps_pots.py",
                           "write_uf3_lammps_pot_files",
                           dict(os=os, datetime=datetime, least_squares=least_squares))
    for tag, path in (("W23", "/root/reference/examples/tungsten_extxyz/model_2and3.json"),
                      ("NeXe", "/root/reference/examples/NeXe_lammps/model_pair.json")):
        model = least_squares.WeightedLinearModel.from_json(path)
        with tempfile.TemporaryDirectory() as tmp:
            writer(chemical_sys=model.bspline_config.chemical_system, model=model, knots_spacing_type="nk",
                   pot_dir=tmp, uf3_lammps_pot_name="pot.uf3", author="golden", lammps_units="metal")
            text = open(os.path.join(tmp, "pot.uf3")).read()
        open(os.path.join(out, f"lammps_{tag}.uf3.txt"), "w").write(text)
        print(tag, len(text), "bytes")
    table = function_from("/root/reference/uf3/forcefield/lammps.py", "export_tabulated_potential",
                          dict(np=np, datetime=datetime, interpolate=interpolate, Tuple=tuple))   # symbols are passed as str
    model = least_squares.WeightedLinearModel.from_json("/root/reference/examples/tungsten_extxyz/model_2and3.json")
    basis = model.bspline_config
    pair = basis.interactions_map[2][0]
    start = basis.get_interaction_partitions()[1][pair]
    size = basis.get_interaction_partitions()[0][pair]
    text = table(basis.knots_map[pair], model.coefficients[start:start + size], pair, grid=200,
                 contributor="golden", rounding=8)
    open(os.path.join(out, "lammps_W23_pair_table.txt"), "w").write(text)
    print("table", len(text), "bytes")


if __name__ == "__main__":
    main()
