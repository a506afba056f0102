"""Pins the CPU oracle (oracle/uf3_oracle.c) against golden vectors produced by the running,
unmodified reference (oracle/make_golden.py): neighbour lists exactly, feature rows and
energies / forces to round-off.  Runs on the CPU."""
import json

import numpy as np
import pytest

import golden_util as gu
from oracle import uf3_oracle as orc


@pytest.mark.parametrize("name", gu.case_names("featurize"))
def test_oracle_features_and_lists(name):
    case = gu.Case(name)
    basis = case.basis()
    assert json.loads(str(case["columns"])) == basis.get_column_names()
    packed = orc.PackedBasis(basis)
    offsets = case.image_offsets(basis)
    n = len(case.numbers)
    if "n_sup" in case:
        assert int(case["n_sup"]) == n * len(offsets)
    want_forces = "x_forces" in case
    xe, xf = orc.featurize(packed, case.positions, case.numbers, offsets, forces=want_forces)
    assert gu.rel_err(xe, case["x_energy"]) < 1e-12
    if want_forces:
        assert gu.rel_err(xf, case["x_forces"]) < 1e-12
    if "nl2_i" in case:
        off, idx = orc.neighbor_lists(packed, case.positions, case.numbers, offsets, 2)
        want_off, want_idx = gu.csr_from_pairs(case["nl2_i"], case["nl2_j"], n)
        assert np.array_equal(off, want_off) and np.array_equal(idx, want_idx)
        if basis.degree > 2:
            off, idx = orc.neighbor_lists(packed, case.positions, case.numbers, offsets, 3)
            want_off, want_idx = gu.csr_from_pairs(case["nl3_i"], case["nl3_j"], n)
            assert np.array_equal(off, want_off) and np.array_equal(idx, want_idx)


@pytest.mark.parametrize("name", gu.case_names("calculator"))
def test_oracle_energy_forces(name):
    case = gu.Case(name)
    basis = case.basis()
    packed = orc.PackedBasis(basis)
    e, f = orc.energy_forces(basis, packed, case["coefficients"], case.positions, case.numbers,
                             case.image_offsets(basis))
    assert abs(e - float(case["energy"])) <= 1e-12 * max(1.0, abs(float(case["energy"])))
    assert gu.rel_err(f, case["forces"]) < 1e-12


def test_reference_known_answers_are_in_the_fixtures():
    """Values printed in the reference's tests/test_calculator.py:40-50, :64-70, :87-98, :109-114."""
    assert np.isclose(float(gu.Case("calc_w_dimer_free")["energy"]), -1.21578, atol=1e-5)
    assert np.allclose(np.abs(gu.Case("calc_w_dimer_free")["forces"]), 3.96244881, atol=1e-6)
    assert np.isclose(float(gu.Case("calc_w_dimer_pbc")["energy"]), -15.33335, atol=1e-5)
    assert np.isclose(float(gu.Case("calc_w_trimer")["energy"]), -18.79979353611411, rtol=1e-12)
    assert np.isclose(float(gu.Case("calc_w8_pbc")["energy"]), -76.358888229785, rtol=1e-11)
    assert np.isclose(float(gu.Case("calc_nexe_dimer")["energy"]), 0.3464031387757268, rtol=1e-12)
