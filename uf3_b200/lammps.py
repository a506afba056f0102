"""Potential hand-off to LAMMPS: the two text formats the reference
writes from a fitted model.

* `write_uf3_lammps_pot_files` — the `.uf3` file read by `pair_style uf3`, as
  `/root/reference/lammps_plugin/scripts/generate_uf3_lammps_pots.py:58-165` lays it out
  (one block per pair and per trio: header, knots, coefficients; 3-body blocks list the
  knot vectors in the order n, m, l and the DECOMPRESSED coefficient grid row by row).
  The reference prints `leading_trim` / `trailing_trim` verbatim; since those became
  per-degree dicts (`bspline.py:66-67`) the file then carries `{2: 0, 3: 3}` where
  `pair_uf3` expects two integers.  Here the block's own degree is looked
  up (`legacy_trim_field=True` reproduces the reference's text byte for byte).
* `export_tabulated_potential` — `pair_style table` input sampled from a pair spline
  (`/root/reference/uf3/forcefield/lammps.py:218-271`); energies and forces carry the
  factor 2 of the reference (LAMMPS does not double-count bonds).  Values are evaluated with
  the same cubic pieces the CUDA kernels use (`uf3b_host_eval_basis`).

Host-side text only: nothing here touches the GPU.
"""
import os
from datetime import datetime

import numpy as np

from uf3_b200 import least_squares
from uf3_b200.elements import chemical_symbols


def _trim(value, degree, legacy):
    if legacy or not isinstance(value, dict):
        return str(value)
    return str(int(value[degree]))


def uf3_lammps_pot_text(model, knots_spacing_type="nk", author="", lammps_units="metal",
                        legacy_trim_field=False, now=None):
    """The `.uf3` potential file of `model` (a `WeightedLinearModel`) as one string."""
    if knots_spacing_type not in ("uk", "nk"):
        raise ValueError(f"Supplied knot spacing type {knots_spacing_type}\n"
                         "is not a valid choice. Only uk or nk are valid types")
    basis = model.bspline_config
    stamp = (now or datetime.now()).strftime("%Y-%m-%d %H:%M:%S")
    sizes, starts = basis.get_interaction_partitions()[:2]
    fmt = "{:.17g}".format
    blocks = []
    for pair in basis.interactions_map[2]:
        text = f"#UF3 POT UNITS: {lammps_units} DATE: {stamp} AUTHOR: {author} CITATION:\n"
        text += f"2B {pair[0]} {pair[1]} {_trim(basis.leading_trim, 2, legacy_trim_field)} " \
                f"{_trim(basis.trailing_trim, 2, legacy_trim_field)} {knots_spacing_type}\n"
        knots = basis.knots_map[pair]
        text += f"{basis.r_max_map[pair]} {len(knots)}\n"
        text += " ".join(fmt(v) for v in knots) + "\n"
        text += f"{sizes[pair]}\n"
        text += " ".join(fmt(v) for v in model.coefficients[starts[pair]:starts[pair] + sizes[pair]]) + "\n#\n"
        blocks.append(text)
    if 3 in basis.interactions_map:
        solutions = least_squares.arrange_coefficients(model.coefficients, basis)
        for trio in basis.interactions_map[3]:
            text = f"#UF3 POT UNITS: {lammps_units} DATE: {stamp} AUTHOR: {author} CITATION:\n"
            text += f"3B {trio[0]} {trio[1]} {trio[2]} {_trim(basis.leading_trim, 3, legacy_trim_field)} " \
                    f"{_trim(basis.trailing_trim, 3, legacy_trim_field)} {knots_spacing_type}\n"
            r_max, knots = basis.r_max_map[trio], basis.knots_map[trio]
            text += f"{r_max[2]} {r_max[1]} {r_max[0]} {len(knots[2])} {len(knots[1])} {len(knots[0])}\n"
            for leg in (2, 1, 0):
                text += " ".join(fmt(v) for v in knots[leg]) + "\n"
            grid = basis.decompress_3B(solutions[trio], trio)
            text += f"{grid.shape[0]} {grid.shape[1]} {grid.shape[2]}\n"
            for i in range(grid.shape[0]):
                for j in range(grid.shape[1]):
                    text += " ".join(map(str, grid[i, j])) + "\n"
            text += "#\n"
            blocks.append(text)
    return "".join(blocks)


def write_uf3_lammps_pot_files(chemical_sys, model, knots_spacing_type, pot_dir, uf3_lammps_pot_name,
                               author, lammps_units, legacy_trim_field=False):
    """Reference signature (generate_uf3_lammps_pots.py:58-64); returns the file's path."""
    if list(chemical_sys.element_list) != list(model.bspline_config.element_list):
        raise ValueError("chemical system and model disagree on the element list")
    os.makedirs(pot_dir, exist_ok=True)
    path = os.path.join(pot_dir, uf3_lammps_pot_name)
    with open(path, "w") as fh:
        fh.write(uf3_lammps_pot_text(model, knots_spacing_type, author, lammps_units, legacy_trim_field))
    return path


def lammps_input_lines(model, pot_dir, uf3_lammps_pot_name):
    """The two lines the reference script prints for the LAMMPS input (:47-50)."""
    elements = model.bspline_config.element_list
    return (f"pair_style\tuf3 {model.bspline_config.degree} {len(elements)}\n"
            f"pair_coeff\t* * {pot_dir}/{uf3_lammps_pot_name} " + " ".join(elements))


def _spline_and_derivative(knots, coefficients, r):
    """S(r), S'(r) of a clamped cubic spline; zero outside the knot range (BSpline with
    extrapolation would continue the end polynomials; the table only samples inside)."""
    from uf3_b200 import _native
    import ctypes as C
    lib = _native.lib()
    knots = np.ascontiguousarray(knots, dtype=np.float64)
    if r <= knots[3]:           # the kernels' intervals are left-open; the table starts ON the first knot
        r = np.nextafter(knots[3], np.inf)
    v, dv = (C.c_double * 4)(), (C.c_double * 4)()
    idx = lib.uf3b_host_eval_basis(knots.ctypes.data_as(C.POINTER(C.c_double)), len(knots), float(r), v, dv)
    if idx < 0:
        return 0.0, 0.0
    c = coefficients[idx:idx + 4]
    return float(np.dot(c, list(v)[:len(c)])), float(np.dot(c, list(dv)[:len(c)]))


def export_tabulated_potential(knot_sequence, coefficients, interaction, grid=None, filename=None,
                               contributor=None, rounding=6):
    """`pair_style table` text for one pair interaction (lammps.py:218-271)."""
    date = datetime.now().strftime("%m/%d/%Y")
    contributor = contributor or ""
    if not isinstance(interaction[0], str):
        interaction = [chemical_symbols[int(z)] for z in interaction]
    interaction = "-".join(interaction)
    knot_sequence = np.asarray(knot_sequence, dtype=np.float64)
    coefficients = np.asarray(coefficients, dtype=np.float64)
    if grid is None:
        grid = 100
    x_table = np.linspace(knot_sequence[0], knot_sequence[-1], grid) if isinstance(grid, int) else grid
    p_line = "{{0}} {{1:.{0}f}} {{2:.{0}f}} {{3:.{0}f}}".format(rounding)
    lines = ["# DATE: {}  UNITS: metal  CONTRIBUTOR: {}".format(date, contributor),
             "# Ultra-Fast Force Field for {}\n".format(interaction),
             "UF_{}".format(interaction),
             "N {}\n".format(len(x_table))]
    for i, r in enumerate(x_table):
        s, ds = _spline_and_derivative(knot_sequence, coefficients, r)
        lines.append(p_line.format(i + 1, r, s * 2, -ds * 2))
    text = "\n".join(lines)
    if filename is not None:
        with open(filename, "w") as fh:
            fh.write(text)
    return text
