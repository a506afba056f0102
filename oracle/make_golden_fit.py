#!/usr/bin/env python
"""Golden fixture for BASELINE.json configs[0] (featurize + fit plumbing), written by the
UNMODIFIED reference (build container only):

    python oracle/make_golden_fit.py        # writes tests/golden/fit_w2body_plumbing.npz

Eight rattled 128-atom bcc-W frames (4x4x4 cells, a = 3.165 A, sigma = 0.05 A, seeds 0-7) are
labelled with the reference's shipped pair model (examples/tungsten_extxyz/model_2.json)
through its own UFCalculator plus N(0, 1e-3) noise, featurized with its BasisFeaturizer
(2-body basis of pair_potential_demo.ipynb: 0.001-5.5 A, 15 intervals, trailing trim 3) and
fitted with its WeightedLinearModel.fit (regularizer of the notebook: ridge_1b = 1e-16,
ridge_2b = 1e-20 -> raised to 1e-8 here so that the normal equations are well conditioned
and the coefficients are comparable digit by digit; curvature_2b = 1e-6; weight 0.8).
Stored: frames, labels, the reference's coefficient vector and its predictions.

TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(HERE, "ref_shims"), "/root/reference"]
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
warnings.simplefilter("ignore")

import ase  # noqa: E402  (stand-in)
from uf3.data import composition  # noqa: E402
from uf3.representation import bspline, process  # noqa: E402
from uf3.regression import least_squares  # noqa: E402
from uf3.forcefield import calculator  # noqa: E402


def bcc_w(reps, a=3.165, sigma=0.05, seed=0):
    rng = np.random.default_rng(seed)
    base = np.array([[0, 0, 0], [0.5, 0.5, 0.5]]) * a
    cells = np.array([[i, j, k] for i in range(reps[0]) for j in range(reps[1]) for k in range(reps[2])]) * a
    pos = (cells[:, None, :] + base[None, :, :]).reshape(-1, 3)
    pos = pos + rng.normal(0, sigma, pos.shape)
    return ase.Atoms(numbers=[74] * len(pos), positions=pos, cell=np.diag(np.array(reps) * a), pbc=True)


def main():
    label_model = least_squares.WeightedLinearModel.from_json(
        "/root/reference/examples/tungsten_extxyz/model_2.json")
    calc = calculator.UFCalculator(label_model)
    rng = np.random.default_rng(2024)
    rows, frames = {}, []
    for seed in range(8):
        geom = bcc_w((4, 4, 4), seed=seed)
        geom.calc = calc
        energy = geom.get_potential_energy() + rng.normal(0, 1e-3)
        forces = geom.get_forces() + rng.normal(0, 1e-3, (len(geom), 3))
        geom.calc = None
        rows[f"w_{seed}"] = dict(geometry=geom, energy=energy, fx=forces[:, 0], fy=forces[:, 1], fz=forces[:, 2])
        frames.append((geom.get_positions(), np.array(geom.get_cell()), energy, forces))
    df_data = pd.DataFrame.from_dict(rows, orient="index")

    chem = composition.ChemicalSystem(["W"], degree=2)
    basis = bspline.BSplineBasis(chem, r_min_map={("W", "W"): 0.001}, r_max_map={("W", "W"): 5.5},
                                 resolution_map={("W", "W"): 15}, trailing_trim=3)
    featurizer = process.BasisFeaturizer(basis)
    df_features = featurizer.evaluate(df_data, progress=None)
    x_e, y_e, x_f, y_f = least_squares.dataframe_to_tuples(df_features, n_elements=1, energy_key="energy")
    strengths = dict(ridge_1b=1e-16, ridge_2b=1e-8, curvature_2b=1e-6)
    model = least_squares.WeightedLinearModel(basis, **strengths)
    model.fit(x_e, y_e, x_f, y_f, weight=0.8)
    out = os.path.join(REPO, "tests", "golden", "fit_w2body_plumbing.npz")
    np.savez_compressed(
        out, positions=np.stack([f[0] for f in frames]), cell=frames[0][1],
        energies=np.array([f[2] for f in frames]), forces=np.stack([f[3] for f in frames]),
        coefficients=np.asarray(model.coefficients), predict_e=model.predict(x_e), predict_f=model.predict(x_f),
        x_e=x_e, y_e=y_e, n_force_rows=len(y_f), ridge_1b=1e-16, ridge_2b=1e-8, curvature_2b=1e-6, weight=0.8)
    print("wrote", out, os.path.getsize(out) // 1024, "KiB; coefficients", model.coefficients[:5])


if __name__ == "__main__":
    main()
