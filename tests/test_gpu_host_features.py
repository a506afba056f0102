"""GPU tests of the callers either side of the hot path that round 2 added: frames dealt over the
visible GPUs (`evaluate_parallel`, process.py:196-254), the chunked feature store with its
`fit_from_file` reader (process.py:256-291, least_squares.py:355-433) and `UFCalculator.relax_fmax`
(calculator.py:406-436)."""
import warnings
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pandas as pd
import pytest

import golden_util as gu
from uf3_b200 import least_squares as ls, store, synthetic
from uf3_b200.atoms import Atoms
from uf3_b200.calculator import UFCalculator
from uf3_b200.process import BasisFeaturizer

pytestmark = pytest.mark.gpu


def _w_model():
    case = gu.Case("calc_syn_w54_model23")
    model = ls.WeightedLinearModel(case.basis())
    model.coefficients = np.array(case["coefficients"])
    return case, model


def _frames(n_frames, cells=(2, 2, 2), seed0=0):
    case, model = _w_model()
    calc = UFCalculator(model)
    rows = {}
    for k in range(n_frames):
        pos, numbers, cell, pbc = synthetic.bcc_w(cells, a=3.17 + 0.01 * k, sigma=0.08, seed=seed0 + k)
        atoms = Atoms(numbers=numbers, positions=pos, cell=cell, pbc=pbc)
        f = calc.get_forces(atoms)
        rows[f"w_{k}"] = {"geometry": atoms, "energy": calc.get_potential_energy(atoms),
                          "fx": f[:, 0], "fy": f[:, 1], "fz": f[:, 2]}
    return pd.DataFrame.from_dict(rows, orient="index"), synthetic.w_basis("demo")


def test_evaluate_parallel_equals_evaluate():
    df, basis = _frames(5)
    feat = BasisFeaturizer(basis)
    serial = feat.evaluate(df)
    parallel = feat.evaluate_parallel(df, n_jobs=3, shuffle=True)
    assert list(parallel.index) == list(serial.index)
    assert np.array_equal(parallel.to_numpy(), serial.to_numpy())
    with ThreadPoolExecutor(max_workers=2) as pool:       # the reference passes an executor as `client`
        pooled = feat.evaluate_parallel(df, pool, n_jobs=2, shuffle=False)
    assert np.array_equal(pooled.to_numpy(), serial.to_numpy())
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        single = feat.evaluate_parallel(df, n_jobs=1)
    assert any("serial" in str(w.message) for w in caught)
    assert np.array_equal(single.to_numpy(), serial.to_numpy())


def test_feature_store_and_fit_from_file_on_the_device(tmp_path):
    df, basis = _frames(6)
    feat = BasisFeaturizer(basis)
    path = str(tmp_path / "features.h5")
    feat.batched_to_hdf(path, df.iloc[:2], n_jobs=2, batch_size=2)        # interrupted after one table
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        feat.batched_to_hdf(path, df, n_jobs=2, batch_size=2)
    n_chunks, n_entries, names, _ = store.analyze_hdf_tables(path)
    n_atoms = len(df.iloc[0]["geometry"])
    assert n_chunks == 3 and n_entries == len(df) * (1 + 3 * n_atoms)
    stacked = pd.concat(store.dataframe_batch_loader(path, names))
    direct = feat.evaluate(df)
    assert np.array_equal(stacked.loc[list(df.index)].to_numpy(), direct.to_numpy())

    subset = list(df.index[:5])
    params = dict(ridge_1b=1e-6, ridge_2b=1e-6, ridge_3b=1e-6, curvature_2b=1e-6)
    host = ls.WeightedLinearModel(basis, **params)
    host.fit_from_file(path, subset, weight=0.5, gram="host")
    device = ls.WeightedLinearModel(basis, solver="cusolver", **params)
    device.fit_from_file(path, subset, weight=0.5)                        # device Gram + cuSOLVER
    # the two solves see the same normal equations up to summation order
    x = direct.to_numpy()[:, 1:]
    assert np.allclose(x @ device.coefficients, x @ host.coefficients, rtol=1e-6, atol=1e-6 * np.abs(x @ host.coefficients).max())


def test_relax_fmax_lowers_forces_and_energy():
    _, model = _w_model()
    calc = UFCalculator(model)
    pos, numbers, cell, pbc = synthetic.bcc_w((2, 2, 2), a=3.25, sigma=0.06, seed=3)
    atoms = Atoms(numbers=numbers, positions=pos, cell=cell, pbc=pbc)
    e0 = calc.get_potential_energy(atoms)
    f0 = np.abs(calc.get_forces(atoms)).max()
    fixed_cell = calc.relax_fmax(atoms, fmax=0.02, relax_cell=False, timeout=120.0)
    assert np.allclose(fixed_cell.get_cell(), cell)
    assert np.sqrt((calc.get_forces(fixed_cell) ** 2).sum(axis=1).max()) < 0.02 < f0
    assert calc.get_potential_energy(fixed_cell) < e0
    relaxed = calc.relax_fmax(atoms, fmax=0.02, relax_cell=True, timeout=120.0)
    assert np.sqrt((calc.get_forces(relaxed) ** 2).sum(axis=1).max()) < 0.02
    assert calc.get_potential_energy(relaxed) <= calc.get_potential_energy(fixed_cell) + 1e-9
    stress = calc.get_stress(relaxed)
    assert np.abs(stress).max() * relaxed.get_volume() / len(relaxed) < 0.02
    assert np.allclose(atoms.get_positions(), pos)                         # the input is left alone
