"""Host-side mirror of the reference interface (CPU): basis bookkeeping, packing of the
device tables, image tables, model files.  Expected values come from the reference's own
tests (file:line in each test) or from fixtures written by the running reference."""
import json
import pickle

import numpy as np
import pytest

import golden_util as gu
from oracle import uf3_oracle as orc
from uf3_b200 import bspline, composition, geometry, least_squares, synthetic
from uf3_b200.atoms import Atoms
from uf3_b200.tables import BasisTables, bin_map


def test_column_names_match_reference_fixtures():
    for name in gu.case_names("featurize"):
        case = gu.Case(name)
        assert json.loads(str(case["columns"])) == case.basis().get_column_names()


def test_headline_basis_sizes():
    """demo basis F = 73 (1 + 18 + 54), manuscript F = 456, Ne-Xe F = 56."""
    demo, manuscript, nexe = synthetic.w_basis("demo"), synthetic.w_basis("manuscript"), synthetic.nexe_basis()
    assert (demo.n_feats, manuscript.n_feats, nexe.n_feats) == (73, 456, 56)
    assert demo.partition_sizes == [1, 18, 54]
    assert demo.symmetry[("W", "W", "W")] == 2
    assert demo.r_cut == 5.5 and nexe.r_cut == 8.0
    assert list(demo.col_idx) == [18, 17, 16]       # trailing trim of the pair block


def test_symmetry_classes():
    """tests/test_bsplines.py:117-245 of the reference (representative cases)."""
    f = bspline.find_symmetry_3B
    assert f(("W", "W", "W"), [1, 1, 1], [4, 4, 4], [5, 5, 5]) == 3
    assert f(("W", "W", "W"), [1, 1, 1], [4, 4, 8], [5, 5, 10]) == 2
    assert f(("W", "W", "W"), [1, 1, 1], [4, 5, 8], [5, 5, 10]) == 1
    assert f(("A", "B", "B"), [1, 1, 1], [4, 4, 4], [5, 5, 5]) == 2
    assert f(("A", "B", "C"), [1, 1, 1], [4, 4, 4], [5, 5, 5]) == 1


def test_uniform_knots_and_defaults():
    """bspline.py:243-258, :1032-1035 of the reference."""
    chem = composition.ChemicalSystem(["Ar"], degree=3)
    basis = bspline.BSplineBasis(chem)
    assert basis.r_min_map[("Ar", "Ar")] == 1.0 and basis.r_max_map[("Ar", "Ar")] == 8.0
    assert basis.resolution_map[("Ar", "Ar", "Ar")] == [5, 5, 10]
    assert basis.r_max_map[("Ar", "Ar", "Ar")] == [4.0, 4.0, 8.0]
    knots = basis.knots_map[("Ar", "Ar")]
    assert len(knots) == 15 + 7 and np.all(knots[:4] == 1.0) and np.all(knots[-4:] == 8.0)
    assert basis.leading_trim == {2: 0, 3: 3} and basis.trailing_trim == {2: 3, 3: 3}


def test_supercell_factors_and_image_order():
    """tests/test_geometry.py:15-30 of the reference: factors [3, 2, 1]; sizes 54 / 210."""
    cell = np.array([[2.0, 0, 0], [3.0, 1.5, 0], [0.5, 0, 2.5]])
    # per-axis order [0, 1, -1, 2, -2, ...] and rank 0 = home cell
    abc, offsets = geometry.image_table(cell, [True, True, True], 3.0)
    assert tuple(abc[0]) == (0, 0, 0) and np.all(offsets[0] == 0)
    table = {tuple(row) for row in abc}
    assert all(tuple(-np.array(row)) in table for row in table)      # inversion symmetric
    assert np.allclose(offsets, abc @ cell)
    abc2, _ = geometry.image_table(cell, [True, False, True], 3.0)
    assert np.all(abc2[:, 1] == 0)
    atoms = Atoms("Au2", positions=[[0, 0, 0], [0.5, 0.3, 0.2]], cell=cell, pbc=True)
    sup = geometry.get_supercell(atoms, r_cut=3.0)
    assert len(sup) == 2 * len(abc)
    assert np.array_equal(sup.get_positions()[:2], atoms.get_positions())


@pytest.mark.parametrize("name", ["syn_w54_demo", "ref_steel_pbc", "syn_ternary_triclinic", "ref_h2o_trimB"])
def test_bin_map_equals_compress_3b(name):
    """tables.bin_map restates compress_3B (bspline.py:664-690); the oracle probes the same
    function by brute force with unit grids."""
    basis = gu.Case(name).basis()
    tables = BasisTables(basis)
    packed = orc.PackedBasis(basis)
    assert np.array_equal(tables.bin_col, packed.bin_col)
    assert np.array_equal(tables.bin_weight, packed.bin_w)
    rng = np.random.default_rng(0)
    for t, trio in enumerate(tables.trios):
        shape = tuple(len(k) - 4 for k in basis.knots_map[trio])
        grid = rng.normal(size=shape)
        col, w = bin_map(basis, trio)
        folded = np.zeros(len(basis.template_mask[trio]))
        np.add.at(folded, col[col >= 0], (grid.ravel() * w)[col >= 0])
        assert np.allclose(folded, basis.compress_3B(grid, trio), rtol=1e-13, atol=1e-13)
        # decompress_3B through the same map (what uf3b_basis_set_coefficients builds)
        coeff = rng.normal(size=basis.n_feats)
        start, size = tables.partition[trio]
        want = basis.decompress_3B(coeff[start:start + size], trio).ravel()
        assert np.allclose(tables.decompressed_grid(coeff, t), want, rtol=1e-13, atol=1e-13)


def test_symmetry3_bin_map():
    chem = composition.ChemicalSystem(["Si"], degree=3)
    basis = bspline.BSplineBasis(chem, r_min_map={("Si", "Si", "Si"): [1.0, 1.0, 1.0]},
                                 r_max_map={("Si", "Si", "Si"): [4.0, 4.0, 4.0]},
                                 resolution_map={("Si", "Si", "Si"): [5, 5, 5]},
                                 leading_trim={2: 0, 3: 0}, trailing_trim={2: 3, 3: 1})
    assert basis.symmetry[("Si", "Si", "Si")] == 3
    tables = BasisTables(basis)
    packed = orc.PackedBasis(basis)
    assert np.array_equal(tables.bin_col, packed.bin_col)
    assert np.allclose(tables.bin_weight, packed.bin_w, rtol=0, atol=1e-15)
    # 2 x 1/2, 6 x 1/6: every kept bin folds with weight one (up to round-off)
    assert np.allclose(tables.bin_weight[tables.bin_col >= 0], 1.0, rtol=0, atol=1e-15)


def test_basis_pickles_without_scipy_callables():
    basis = synthetic.w_basis("demo")
    _ = basis.basis_functions
    clone = pickle.loads(pickle.dumps(basis))
    assert clone.get_column_names() == basis.get_column_names()
    assert clone._basis_functions is None


def test_model_round_trip(tmp_path):
    case = gu.Case("calc_syn_w54_model23")
    basis = case.basis()
    model = least_squares.WeightedLinearModel(basis)
    model.coefficients = np.array(case["coefficients"])
    path = tmp_path / "model.json"
    model.to_json(str(path))
    loaded = least_squares.WeightedLinearModel.from_json(str(path))
    assert np.allclose(loaded.coefficients, model.coefficients, rtol=1e-13, atol=1e-15)
    assert loaded.bspline_config.get_column_names() == basis.get_column_names()
    with pytest.raises(ValueError):
        least_squares.WeightedLinearModel(basis, data_coverage=np.zeros(3))


def test_trim_validation_and_unknown_strategy():
    chem = composition.ChemicalSystem(["W"])
    with pytest.raises(ValueError):
        bspline.BSplineBasis(chem, leading_trim={"2": 0})
    with pytest.raises(ValueError):
        bspline.BSplineBasis(chem, knot_strategy="cubic")


def test_fit_spline_1d_reproduces_the_reference_fit():
    """bspline.fit_spline_1d (bspline.py:898-950): the Lennard-Jones curve the reference's calculator test
    fits (tests/test_calculator.py:20-33); its coefficients are in the fixture the reference wrote."""
    import golden_util as gu
    from uf3_b200 import bspline
    case = gu.Case("calc_w_dimer_free")
    knots = np.array(case.basis().knots_map[("W", "W")])
    x = np.linspace(2.0, 6.0, 1000)
    y = 4 * 0.87 * ((2.5 / x) ** 12 - (2.5 / x) ** 6)
    coefficients = bspline.fit_spline_1d(x, y, knots)
    want = np.array(case["coefficients"])[1:]
    assert np.abs(coefficients - want).max() <= 1e-10 * np.abs(want).max()
