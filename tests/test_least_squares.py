"""Normal-equation bookkeeping (CPU): frozen columns, E/F weighting, and the flattened
Gram statistics that multi-GPU fits all-reduce.  The reference behaviour restated here is
regression/least_squares.py:248-353, :666-771, :817-890, :1147-1168."""
import numpy as np
import pytest

import golden_util as gu
from uf3_b200 import least_squares as ls


def _rows(name="syn_w54_demo", seed=0):
    case = gu.Case(name)
    rng = np.random.default_rng(seed)
    coeff = rng.normal(size=case["x_energy"].shape[0])
    x_e = np.stack([case["x_energy"], case["x_energy"] * 1.01 + 0.1])
    n_atoms = len(case.numbers)
    x_f = case["x_forces"]
    y_e = x_e @ coeff + rng.normal(0, 1e-3, len(x_e))
    y_f = x_f @ coeff + rng.normal(0, 1e-3, len(x_f))
    return case, n_atoms, x_e, y_e, x_f, y_f


def test_freeze_and_revert_columns():
    x = np.arange(20.0).reshape(4, 5)
    y = np.arange(4.0)
    mask = ls.get_freezing_mask(5, np.array([1, 3]))
    assert mask.tolist() == [0, 2, 4]
    xm, ym = ls.freeze_columns(x, y, mask, np.array([2.0, -1.0]), np.array([1, 3]))
    assert xm.shape == (4, 3) and np.allclose(ym, y - 2 * x[:, 1] + x[:, 3])
    full = ls.revert_frozen_coefficients(np.array([1.0, 2.0, 3.0]), 5, mask, np.array([9.0, 8.0]), np.array([1, 3]))
    assert full.tolist() == [1, 9, 2, 8, 3]


def test_e_f_weights():
    assert ls.calc_E_F_weights(4, 9, 0.0, 2.0) == (1.0, 1 / 3)
    w_e, w_f = ls.calc_E_F_weights(4, 9, 0.5, 2.0)
    assert np.isclose(w_e, 1 / 2 / 0.5) and np.isclose(w_f, 1 / 3 / 2.0)


def test_fit_from_statistics_equals_fit_from_rows():
    case, n_atoms, x_e, y_e, x_f, y_f = _rows()
    basis = case.basis()
    direct = ls.WeightedLinearModel(basis, ridge_1b=1e-4, ridge_2b=1e-4, ridge_3b=1e-4, curvature_2b=1e-4)
    # dataframe_to_tuples divides energy rows and targets by the atom count (:697-700)
    direct.fit(x_e / n_atoms, y_e / n_atoms, x_f, y_f, weight=0.7)
    stats = ls.GramStats(basis.n_feats)
    for row, target in zip(x_e, y_e):
        stats.add_energy_row(row, target, n_atoms)
    half = len(x_f) // 2
    stats.add_force_rows(x_f[:half], y_f[:half])
    stats.add_force_rows(x_f[half:], y_f[half:])
    via_stats = ls.WeightedLinearModel(basis, ridge_1b=1e-4, ridge_2b=1e-4, ridge_3b=1e-4, curvature_2b=1e-4)
    via_stats.fit_from_accumulator(stats, weight=0.7)
    assert np.allclose(via_stats.coefficients, direct.coefficients, rtol=1e-6, atol=1e-7)
    assert np.all(via_stats.coefficients[basis.col_idx] == 0)
    # flatten / restore is loss-free
    clone = ls.GramStats(basis.n_feats)
    clone.from_vector(stats.to_vector())
    assert np.array_equal(clone.to_vector(), stats.to_vector())
    assert len(stats.to_vector()) == 2 * basis.n_feats ** 2 + 2 * basis.n_feats + 6


def test_energy_only_fit():
    case, n_atoms, x_e, y_e, _, _ = _rows()
    basis = case.basis()
    model = ls.WeightedLinearModel(basis, ridge_2b=1e-6, ridge_3b=1e-6, ridge_1b=1e-6)
    model.fit(x_e / n_atoms, y_e / n_atoms)
    assert model.coefficients.shape == (basis.n_feats,)
    assert np.isfinite(model.coefficients).all()


def test_fit_reproduces_the_reference_fit_on_its_own_rows():
    """BASELINE.json configs[0] plumbing, regression half: WeightedLinearModel.fit on the
    reference's energy rows of tests/golden/fit_w2body_plumbing.npz (written by the running
    reference, oracle/make_golden_fit.py) — energy-only here because the fixture does not
    carry the 3072 force rows; the full pipeline is compared on the GPU
    (tests/test_gpu_api.py::test_plumbing_featurize_and_fit_matches_reference)."""
    import os
    from uf3_b200 import bspline, composition
    fix = np.load(os.path.join(os.path.dirname(__file__), "golden", "fit_w2body_plumbing.npz"))
    chem = composition.ChemicalSystem(["W"], degree=2)
    basis = bspline.BSplineBasis(chem, r_min_map={("W", "W"): 0.001}, r_max_map={("W", "W"): 5.5},
                                 resolution_map={("W", "W"): 15}, trailing_trim=3)
    assert fix["x_e"].shape == (8, basis.n_feats)
    # energy rows are per atom: the composition column is 1 after the division by 128
    assert np.allclose(fix["x_e"][:, 0], 1.0)
    model = ls.WeightedLinearModel(basis, ridge_1b=float(fix["ridge_1b"]), ridge_2b=float(fix["ridge_2b"]),
                                   curvature_2b=float(fix["curvature_2b"]))
    model.fit(fix["x_e"], fix["y_e"])
    assert np.isfinite(model.coefficients).all() and np.all(model.coefficients[basis.col_idx] == 0)
    # the reference model (fitted on energies AND forces) predicts the same energies to the noise level
    assert np.abs(model.predict(fix["x_e"]) - fix["predict_e"]).max() < 5e-3


def test_variance_recorder_and_gram_from_df_reproduce_fit():
    """The reference's streaming pieces (least_squares.py:19-68, :425-483): Gram blocks per chunk + running
    target statistics give the coefficients of the one-shot fit."""
    import pandas as pd
    case, n_atoms, x_e, y_e, x_f, y_f = _rows()
    basis = case.basis()
    rec = ls.VarianceRecorder()
    for part in np.array_split(y_f, 5):
        rec.update(part)
    assert rec.n == len(y_f) and np.isclose(rec.mean, y_f.mean()) and np.isclose(rec.std, y_f.std())
    # two configurations as a feature frame (target first, then the row)
    rows, index = [], []
    half = len(x_f) // 2
    for k, name in enumerate(("a", "b")):
        rows.append(np.insert(x_e[k], 0, y_e[k])); index.append((name, "energy"))
        for j, (xr, yr) in enumerate(zip(x_f[k * half:(k + 1) * half], y_f[k * half:(k + 1) * half])):
            rows.append(np.insert(xr, 0, yr)); index.append((name, f"fx_{j}"))
    df = pd.DataFrame(np.array(rows), index=pd.MultiIndex.from_tuples(index))
    params = dict(ridge_1b=1e-3, ridge_2b=1e-3, ridge_3b=1e-3, curvature_2b=1e-3)
    model = ls.WeightedLinearModel(basis, **params)
    gram_e, gram_f, ord_e, ord_f = model.initialize_gram_ordinate()
    e_var, f_var = ls.VarianceRecorder(), ls.VarianceRecorder()
    for name in ("a", "b"):
        g_e, g_f, o_e, o_f = model.gram_from_df(df, [name], e_variance=e_var, f_variance=f_var)
        gram_e += g_e; gram_f += g_f; ord_e += o_e; ord_f += o_f
    w_e, w_f = ls.calc_E_F_weights(e_var.n, f_var.n, e_var.std, f_var.std)
    model.fit_with_gram(*model.combine_weighted_gram(gram_e, gram_f, ord_e, ord_f, w_e, w_f, 0.4))
    direct = ls.WeightedLinearModel(basis, **params)
    t_e, u_e, t_f, u_f = ls.dataframe_to_tuples(df, n_elements=1)
    direct.fit(t_e, u_e, t_f, u_f, weight=0.4)
    assert np.allclose(model.coefficients, direct.coefficients, rtol=1e-6, atol=1e-9)
    y1, p1, y2, p2 = ls.subset_prediction(df, model, subset_keys=["b"], n_elements=1)
    assert len(y1) == 1 and len(y2) == half and np.allclose(p2, t_f[half:] @ model.coefficients)
    sol = ls.weighted_least_squares(x_f[:, 1:19], y_f, weights=np.ones(len(y_f)), regularizer=1e-3 * np.eye(18))
    assert sol.shape == (18,) and np.all(np.isfinite(sol))
    with pytest.raises(ValueError):
        ls.validate_regularizer(np.eye(4), 5)
