"""Chunked feature store (uf3_b200/store.py), `BasisFeaturizer.batched_to_hdf` resume semantics and
`WeightedLinearModel.fit_from_file` — reference: representation/process.py:256-291, :538-562,
data/io.py:943-970, regression/least_squares.py:355-433."""
import os
import warnings

import numpy as np
import pandas as pd
import pytest

import golden_util as gu
from uf3_b200 import least_squares as ls
from uf3_b200 import process, store


def _feature_frame(case, names, seed=0):
    """Rows in the layout `evaluate` returns: per configuration an energy row and 3N force rows,
    target in the first column."""
    rng = np.random.default_rng(seed)
    basis = case.basis()
    coeff = rng.normal(size=basis.n_feats) * 0.1
    n = len(case.numbers)
    rows, index = [], []
    for k, name in enumerate(names):
        x_e = case["x_energy"] * (1.0 + 0.01 * k)
        x_f = case["x_forces"] * (1.0 - 0.02 * k)
        rows.append(np.insert(x_e, 0, x_e @ coeff + 0.05 * k * n))       # spread per-atom energies
        index.append((name, "energy"))
        y_f = x_f @ coeff + rng.normal(0, 1e-3, len(x_f))
        for j, comp in enumerate("xyz"):
            for i in range(n):
                rows.append(np.insert(x_f[j * n + i], 0, y_f[j * n + i]))
                index.append((name, f"f{comp}_{i}"))
    columns = basis.get_column_names()
    return pd.DataFrame(np.array(rows), index=pd.MultiIndex.from_tuples(index), columns=columns)


def test_store_round_trip_and_analysis(tmp_path):
    case = gu.Case("syn_w54_demo")
    df = _feature_frame(case, ["a", "b", "c"])
    path = str(tmp_path / "features.h5")
    first, second = df.loc[["a", "b"]], df.loc[["c"]]
    store.save_feature_db(first, path, table_name="features_000")
    store.save_feature_db(second, path, table_name="features_001")
    n_chunks, n_entries, names, lengths = store.analyze_hdf_tables(path)
    assert (n_chunks, n_entries, names) == (2, len(df), ["features_000", "features_001"])
    assert lengths == {"features_000": len(first), "features_001": len(second)}
    back = store.load_feature_db(path, "features_001")
    assert list(back.columns) == list(df.columns)
    assert list(back.index) == list(second.index)
    assert np.array_equal(back.to_numpy(), second.to_numpy())
    loaded = list(store.dataframe_batch_loader(path, names))
    assert np.array_equal(pd.concat(loaded).to_numpy(), df.to_numpy())
    if not store.have_pytables():
        with pytest.raises(ValueError):
            store.save_feature_db(second, path, table_name="features_001")


class _CountingFeaturizer(process.BasisFeaturizer):
    """`evaluate` served from prepared rows, so that the store logic runs without a GPU."""

    def __init__(self, basis, frame):
        super().__init__(basis)
        self.frame, self.calls = frame, []

    def evaluate(self, df_data, **kwargs):
        self.calls.append(list(df_data.index))
        return self.frame.loc[list(df_data.index)]

    def evaluate_parallel(self, df_data, client=None, **kwargs):
        return self.evaluate(df_data)


def test_batched_store_resumes_and_fit_from_file_matches_fit(tmp_path):
    case = gu.Case("syn_w54_demo")
    basis = case.basis()
    names = [f"cfg_{k}" for k in range(7)]
    frame = _feature_frame(case, names)
    df_data = pd.DataFrame({"geometry": [None] * len(names)}, index=names)
    path = str(tmp_path / "store.h5")
    feat = _CountingFeaturizer(basis, frame)
    feat.batched_to_hdf(path, df_data.iloc[:4], batch_size=2)            # an interrupted run: 2 of 4 tables
    assert store.analyze_hdf_tables(path)[2] == ["features_000", "features_001"]
    feat.calls.clear()
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        feat.batched_to_hdf(path, df_data, batch_size=2)
    assert any("already exists: contains 2 chunks" in str(w.message) for w in caught)
    assert feat.calls == [names[4:6], names[6:7]]                           # existing tables are skipped
    n_chunks, n_entries, _, _ = store.analyze_hdf_tables(path)
    assert (n_chunks, n_entries) == (4, len(frame))

    subset = names[:5]
    params = dict(ridge_1b=1e-2, ridge_2b=1e-2, ridge_3b=1e-2, curvature_2b=1e-2)     # well conditioned: the frames are near copies
    direct = ls.WeightedLinearModel(basis, **params)
    x_e, y_e, x_f, y_f = ls.dataframe_to_tuples(frame.loc[subset], n_elements=len(basis.element_list))
    direct.fit(x_e, y_e, x_f, y_f, weight=0.3)
    from_file = ls.WeightedLinearModel(basis, **params)
    from_file.fit_from_file(path, subset, weight=0.3, gram="host")
    assert np.allclose(from_file.coefficients, direct.coefficients, rtol=1e-6, atol=1e-8)
    y_e2, p_e, y_f2, p_f, rmse_e, rmse_f = from_file.batched_predict(path, keys=subset)
    assert len(y_e2) == len(subset) and len(y_f2) == len(y_f)
    assert rmse_f < 0.1 * np.std(y_f) and np.allclose(np.sort(y_f2), np.sort(y_f))
    with pytest.raises(FileNotFoundError):
        from_file.fit_from_file(str(tmp_path / "missing.h5"), subset)
    with pytest.raises(ValueError):
        from_file.fit_from_file(path, ["not there"], gram="host")


def test_training_tuples_weights():
    """process.dataframe_to_training_tuples (process.py:574-616): weights per sample class."""
    case = gu.Case("syn_w54_demo")
    df = _feature_frame(case, ["a", "b", "c"])
    x, y, w = process.dataframe_to_training_tuples(df, kappa=0.25)
    is_e = np.asarray(df.index.get_level_values(-1) == "energy")
    assert x.shape == (len(df), df.shape[1] - 1) and np.array_equal(y, df.to_numpy()[:, 0])
    assert np.isclose(w[is_e].sum(), 0.25 / np.std(y[is_e])) and np.isclose(w[~is_e].sum(), 0.75 / np.std(y[~is_e]))
    with pytest.raises(ValueError):
        process.dataframe_to_training_tuples(df, kappa=1.5)
    with pytest.raises(ValueError):
        process.dataframe_to_training_tuples(df.iloc[:1])
