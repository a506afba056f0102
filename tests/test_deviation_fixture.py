"""The one place where the CUDA path knowingly differs from the reference: a trio
whose two neighbours share a species but have different l / m knot vectors (symmetry 1,
bspline.py:723-763).  The fixture (oracle/make_golden_sym1.py, written by the running reference)
holds the reference's energy row at +-delta displacements, so the derivative of ITS OWN energy
features is available without the reference: its force rows are not that derivative (the l / m
assignment by supercell index, angles.py:460-488, flips for ghost-centred triangles), ours are."""
import numpy as np
import pytest

import golden_util as gu

CASE = "dev_w16_sym1"


def _probes(case):
    n = len(case.numbers)
    delta = float(case["probe_delta"])
    fd = -(case["x_energy_plus"] - case["x_energy_minus"]) / (2 * delta)
    rows = [int(axis) * n + int(atom) for atom, axis in zip(case["probe_atoms"], case["probe_axes"])]
    return rows, fd


def test_reference_force_rows_are_not_the_derivative_of_its_energy_row():
    case = gu.Case(CASE)
    assert case.basis().symmetry[("W", "W", "W")] == 1
    rows, fd = _probes(case)
    ref = case["x_forces"][rows]
    n2 = 1 + 18                                   # composition + pair columns: consistent
    assert np.abs(ref[:, :n2] - fd[:, :n2]).max() <= 1e-6 * np.abs(fd[:, :n2]).max()
    assert np.abs(ref[:, n2:] - fd[:, n2:]).max() > 0.05 * np.abs(fd[:, n2:]).max()


def test_oracle_restates_the_reference_on_the_deviating_fixture():
    """The oracle keeps the reference's formulation (ghost-centred triangles ordered by supercell
    index), so it reproduces the reference's rows here; only the CUDA path differs."""
    from oracle import uf3_oracle as orc
    case = gu.Case(CASE)
    basis = case.basis()
    xe, xf = orc.featurize(orc.PackedBasis(basis), case.positions, case.numbers, case.image_offsets(basis))
    assert gu.rel_err(xe, case["x_energy"]) <= 1e-12
    assert gu.rel_err(xf, case["x_forces"]) <= 1e-12


@pytest.mark.gpu
def test_cuda_force_rows_are_the_derivative_of_the_reference_energy_row():
    from uf3_b200 import geometry
    from uf3_b200.engine import Engine
    case = gu.Case(CASE)
    basis = case.basis()
    eng = Engine(basis)
    eng.build_neighbors(case.positions, case.numbers, images=geometry.image_table(case.cell, case.pbc, basis.r_cut))
    xe, xf = eng.featurize()
    eng.close()
    assert gu.rel_err(xe, case["x_energy"]) <= 1e-6                 # energy row: the reference's
    n2 = 1 + 18
    assert gu.rel_err(xf[:, :n2], case["x_forces"][:, :n2]) <= 1e-6  # pair columns: the reference's
    rows, fd = _probes(case)
    assert np.abs(xf[rows] - fd).max() <= 1e-6 * np.abs(fd).max()    # 3-body columns: -d(energy row)/dR
