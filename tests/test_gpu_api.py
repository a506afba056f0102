"""GPU tests of the reference-facing classes (BasisFeaturizer, UFCalculator, GramAccumulator)
against fixtures written by the running reference and against the oracle."""
import pickle
import warnings

import numpy as np
import pytest

import golden_util as gu
from uf3_b200 import geometry, least_squares as ls, synthetic
from uf3_b200.atoms import Atoms
from uf3_b200.calculator import UFCalculator
from uf3_b200.process import BasisFeaturizer

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["ref_steel_pbc", "ref_h2o_trimA", "ref_ch4_trimB", "syn_nexe64_pair"])
def test_evaluate_configuration_rows(name):
    """Row keys, order and values of process.py:293-367 (reference test:
    tests/test_representation.py:605-648 for the Fe-C cell)."""
    case = gu.Case(name)
    feat = BasisFeaturizer(case.basis())
    n = len(case.numbers)
    forces = np.arange(3.0 * n).reshape(3, n)
    rows = feat.evaluate_configuration(case.atoms(), name="cfg", energy=-1.5, forces=forces)
    assert len(rows) == 1 + 3 * n
    keys = list(rows)
    assert keys[0] == ("cfg", "energy") and keys[1] == ("cfg", "fx_0") and keys[-1] == ("cfg", f"fz_{n - 1}")
    assert rows[("cfg", "energy")][0] == -1.5
    assert np.allclose(rows[("cfg", "energy")][1:], case["x_energy"], rtol=1e-5, atol=1e-8)
    got = np.stack([rows[("cfg", f"{c}_{a}")] for c in ("fx", "fy", "fz") for a in range(n)])
    assert np.array_equal(got[:, 0], forces.reshape(-1))
    assert np.allclose(got[:, 1:], case["x_forces"], rtol=1e-5, atol=1e-8)
    assert len(feat.columns) == got.shape[1]
    unnamed = feat.evaluate_configuration(case.atoms(), energy=0.0)
    assert list(unnamed) == ["energy"]


def test_partial_featurizers_and_supercell_argument():
    case = gu.Case("syn_w16_demo")
    basis = case.basis()
    feat = BasisFeaturizer(basis)
    atoms = case.atoms()
    sup = geometry.get_supercell(atoms, r_cut=basis.r_cut)
    n = len(atoms)
    e2, e3 = feat.featurize_energy_2B(atoms, sup), feat.featurize_energy_3B(atoms, sup)
    assert np.allclose(np.concatenate([[n], e2, e3]), case["x_energy"], rtol=1e-5, atol=1e-8)
    f2, f3 = feat.featurize_force_2B(atoms, sup), feat.featurize_force_3B(atoms, sup)
    assert f2.shape == (n, 3, 18) and f3.shape == (n, 3, 54)
    want = case["x_forces"].reshape(3, n, -1).transpose(1, 0, 2)
    assert np.allclose(f2, want[:, :, 1:19], rtol=1e-5, atol=1e-8)
    assert np.allclose(f3, want[:, :, 19:], rtol=1e-5, atol=1e-8)
    # as in the reference, no supercell argument means no periodic images
    isolated = Atoms(numbers=case.numbers, positions=case.positions)
    assert np.allclose(feat.featurize_energy_2B(atoms), feat.featurize_energy_2B(isolated))


def test_unknown_element_warns_and_returns_empty():
    feat = BasisFeaturizer(synthetic.w_basis("demo"))
    geom = Atoms("WFe", positions=[[0, 0, 0], [2.5, 0, 0]])
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        assert feat.evaluate_configuration(geom, name="bad", energy=0.0) == {}
    assert any(issubclass(w.category, RuntimeWarning) for w in caught)


def test_featurizer_pickles_and_dataframe():
    import pandas as pd
    case = gu.Case("syn_w16_demo")
    feat = BasisFeaturizer(case.basis())
    feat.evaluate_configuration(case.atoms(), energy=0.0)
    clone = pickle.loads(pickle.dumps(feat))
    assert clone._engine is None
    n = len(case.numbers)
    df = pd.DataFrame({"geometry": [case.atoms(), case.atoms()], "energy": [1.0, 2.0],
                       "fx": [np.zeros(n)] * 2, "fy": [np.zeros(n)] * 2, "fz": [np.zeros(n)] * 2},
                      index=["a", "b"])
    table = clone.evaluate(df)
    assert table.shape == (2 * (1 + 3 * n), len(clone.columns))
    assert table.index[0] == ("a", "energy")
    assert np.allclose(table.loc[("b", "energy")].to_numpy()[1:], case["x_energy"], rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("name", gu.case_names("calculator"))
def test_calculator_matches_reference(name):
    """Known answers of tests/test_calculator.py:12-114 and synthetic frames."""
    case = gu.Case(name)
    model = ls.WeightedLinearModel(case.basis())
    model.coefficients = np.array(case["coefficients"])
    calc = UFCalculator(model)
    atoms = case.atoms()
    atoms.calc = calc
    assert "W" in calc.solutions or "Ne" in calc.solutions
    assert set(calc.pair_potentials) == set(case.basis().interactions_map[2])
    e = atoms.get_potential_energy()
    f = atoms.get_forces()
    assert abs(e - float(case["energy"])) <= 1e-6 * max(abs(float(case["energy"])), 1e-12)
    assert gu.rel_err(f, case["forces"]) <= 1e-6
    assert calc.results["energy"] == e and calc.results["free_energy"] == e


@pytest.mark.parametrize("name", ["calc_w8_pbc", "calc_syn_w54_model23", "calc_syn_nexe64_pair",
                                  "calc_syn_w128_model23"])
def test_analytic_stress_matches_numerical_stress(name):
    """The kernel's analytic strain derivative against the reference's method (central
    differences over strained cells, calculator.py:399-404) on pair and 2+3-body models.
    (calc_w_dimer_pbc is left out on purpose: its 3 A cell edge puts image pairs exactly on
    r_max = 6 A, where the shipped coefficients are non-zero, so the strained-cell energy
    jumps and the finite difference diverges like 1/d — the analytic value is the limit.)"""
    case = gu.Case(name)
    model = ls.WeightedLinearModel(case.basis())
    model.coefficients = np.array(case["coefficients"])
    atoms = case.atoms()
    analytic = UFCalculator(model)
    atoms.calc = analytic
    got = atoms.get_stress()
    assert got.shape == (6,) and "stress" in analytic.results
    want = UFCalculator(model, numerical_stress=True)._get_stress(atoms, d=1e-5)
    assert np.abs(got - want).max() <= 2e-6 * max(np.abs(want).max(), 1e-3)
    # the 3x3 tensor is symmetric and the virial call leaves energy / forces untouched
    eng = analytic.engine
    e0, f0 = eng.energy_forces()
    e1, f1, w = eng.energy_forces(virial=True)
    assert np.array_equal(w, w.T)
    assert abs(e0 - e1) <= 1e-12 * abs(e0) and gu.rel_err(f1, f0) <= 1e-12


def test_analytic_stress_matches_oracle_energy_differences():
    """Stress from the kernel against central differences of the ORACLE's energy over the six
    strain components (the reference's definition, with the CPU restatement as the energy)."""
    from oracle import uf3_oracle as orc
    from uf3_b200 import geometry
    case = gu.Case("calc_syn_w54_model23")
    basis, coeff = case.basis(), np.array(case["coefficients"])
    model = ls.WeightedLinearModel(basis)
    model.coefficients = coeff
    atoms = case.atoms()
    got = UFCalculator(model)._get_stress(atoms)
    packed = orc.PackedBasis(basis)
    cell0, pos0 = np.array(atoms.get_cell()), np.array(atoms.get_positions())
    numbers, pbc = np.array(atoms.get_atomic_numbers()), np.array(atoms.get_pbc())
    d, want = 1e-5, np.zeros((3, 3))
    for i in range(3):
        for j in range(i, 3):
            energies = []
            for sign in (1, -1):
                strain = np.eye(3)
                if i == j:
                    strain[i, i] += sign * d
                else:
                    strain[i, j] += sign * d / 2
                    strain[j, i] += sign * d / 2
                cell = cell0 @ strain
                pos = pos0 @ strain
                images = geometry.image_table(cell, pbc, basis.r_cut)
                energies.append(orc.energy_forces(basis, packed, coeff, pos, numbers, images[1])[0])
            want[i, j] = want[j, i] = (energies[0] - energies[1]) / (2 * d * abs(np.linalg.det(cell0)))
    want = want.flat[[0, 4, 8, 5, 2, 1]]
    assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()


@pytest.mark.parametrize("name", ["calc_syn_w128_model23", "calc_syn_nexe64_pair", "calc_w8_pbc"])
def test_centre_ranges_sum_to_the_full_frame(name):
    """One frame split into atom ranges (uf3b_neighbors_build_range): the partial energies,
    forces and virials of the ranges add up to the full-frame result, which matches the
    reference fixture; feature rows refuse a partial list."""
    from uf3_b200 import _native, distributed
    from uf3_b200.atoms import frame_arrays
    from uf3_b200.engine import Engine
    case = gu.Case(name)
    basis = case.basis()
    positions, numbers, cell, pbc = frame_arrays(case.atoms())
    images = geometry.image_table(cell, pbc, basis.r_cut)
    eng = Engine(basis)
    eng.set_coefficients(np.array(case["coefficients"]))
    eng.build_neighbors(positions, numbers, images=images)
    e_full, f_full, w_full = eng.energy_forces(virial=True)
    total2 = eng.neighbor_count(2)
    n = len(positions)
    for world in (2, 3):
        e_sum, f_sum, w_sum, pairs = 0.0, np.zeros((n, 3)), np.zeros((3, 3)), 0
        for rank in range(world):
            eng.build_neighbors(positions, numbers, images=images, centres=distributed.atom_range(n, rank, world))
            e, f, w = eng.energy_forces(virial=True)
            e_sum, f_sum, w_sum = e_sum + e, f_sum + f, w_sum + w
            pairs += eng.neighbor_count(2)
            with pytest.raises(_native.UF3BError):
                eng.featurize()
        assert pairs == total2
        assert abs(e_sum - e_full) <= 1e-12 * abs(e_full)
        assert gu.rel_err(f_sum, f_full) <= 1e-12 and gu.rel_err(w_sum, w_full) <= 1e-12
    assert abs(e_full - float(case["energy"])) <= 1e-6 * abs(float(case["energy"]))
    assert gu.rel_err(f_full, case["forces"]) <= 1e-6
    # a rank without centres (more ranks than atoms) contributes exact zeros
    eng.build_neighbors(positions, numbers, images=images, centres=(n, 0))
    e, f, w = eng.energy_forces(virial=True)
    assert e == 0.0 and not f.any() and not w.any() and eng.neighbor_count(2) == 0
    eng.close()


def test_numerical_stress_is_energy_derivative():
    case = gu.Case("calc_w8_pbc")
    model = ls.WeightedLinearModel(case.basis())
    model.coefficients = np.array(case["coefficients"])
    calc = UFCalculator(model, numerical_stress=True)
    atoms = case.atoms()
    stress = calc._get_stress(atoms)
    assert stress.shape == (6,) and np.all(np.isfinite(stress))
    d, vol = 1e-5, atoms.get_volume()
    strained = []
    for sign in (1, -1):
        trial = atoms.copy()
        trial.set_cell(atoms.get_cell() @ np.diag([1 + sign * d, 1, 1]), scale_atoms=True)
        strained.append(calc.get_potential_energy(trial))
    assert np.isclose(stress[0], (strained[0] - strained[1]) / (2 * d * vol), rtol=1e-3, atol=1e-6)


def test_forces_are_minus_energy_gradient():
    case = gu.Case("calc_syn_w54_model23")
    model = ls.WeightedLinearModel(case.basis())
    model.coefficients = np.array(case["coefficients"])
    calc = UFCalculator(model)
    atoms = case.atoms()
    forces = calc.get_forces(atoms)
    h = 1e-5
    for a, c in ((0, 0), (7, 1), (53, 2)):
        plus, minus = atoms.copy(), atoms.copy()
        plus.positions[a, c] += h
        minus.positions[a, c] -= h
        num = -(calc.get_potential_energy(plus) - calc.get_potential_energy(minus)) / (2 * h)
        assert np.isclose(forces[a, c], num, rtol=1e-5, atol=1e-6)
    assert np.abs(forces.sum(axis=0)).max() < 1e-8      # Newton's third law


@pytest.mark.parametrize("kernels", ["cublas", "own"])
def test_device_gram_equals_host_gram_and_fit(kernels, monkeypatch):
    """Both device paths of the normal-equation accumulation (the hand-written k_gram /
    k_ordinate, and cuBLAS dsyrk + dgemv selected by UF3B_GRAM_KERNEL=cublas) against numpy."""
    import torch
    if kernels == "cublas":
        monkeypatch.setenv("UF3B_GRAM_KERNEL", "cublas")
    else:
        monkeypatch.delenv("UF3B_GRAM_KERNEL", raising=False)
    case = gu.Case("syn_w128_demo")
    basis = case.basis()
    feat = BasisFeaturizer(basis)
    n, F = len(case.numbers), basis.n_feats
    eng = feat.engine
    eng.build_neighbors(case.positions, case.numbers, images=geometry.image_table(case.cell, case.pbc, basis.r_cut))
    rows = torch.empty((3 * n, F), dtype=torch.float64, device="cuda")
    xe = np.empty(F)
    eng._featurize_mixed(xe, rows.data_ptr(), F)
    rng = np.random.default_rng(1)
    y_f = rng.normal(size=3 * n)
    dev = ls.GramAccumulator(F)
    dev.add_energy_row(xe, -3.0, n)
    dev.add_force_rows_device(rows.data_ptr(), y_f, 3 * n, F)
    host = ls.GramStats(F)
    host.add_energy_row(case["x_energy"], -3.0, n)
    host.add_force_rows(case["x_forces"], y_f)
    a, b = dev.export(), host.export()
    scale = np.abs(b["gram_f"]).max()
    assert np.abs(a["gram_f"] - b["gram_f"]).max() <= 1e-9 * scale
    assert np.allclose(a["gram_f"], a["gram_f"].T)
    assert np.abs(a["ord_f"] - b["ord_f"]).max() <= 1e-9 * np.abs(b["ord_f"]).max()
    assert np.allclose(a["gram_e"], b["gram_e"], rtol=1e-9, atol=1e-12)
    kw = dict(ridge_1b=1e-4, ridge_2b=1e-4, ridge_3b=1e-4, curvature_2b=1e-4)
    m_dev, m_host = ls.WeightedLinearModel(basis, **kw), ls.WeightedLinearModel(basis, **kw)
    m_dev.fit_from_accumulator(dev)
    m_host.fit_from_accumulator(host)
    assert np.allclose(m_dev.coefficients, m_host.coefficients, rtol=1e-6, atol=1e-7)
    dev.close()


def test_plumbing_featurize_and_fit_matches_reference():
    """BASELINE.json configs[0]: eight 128-atom W frames -> BasisFeaturizer.evaluate (CUDA rows)
    -> dataframe_to_tuples -> WeightedLinearModel.fit, against the coefficient vector the
    running reference obtained from the same frames and labels (oracle/make_golden_fit.py)."""
    import os
    import pandas as pd
    from uf3_b200 import bspline, composition
    fix = np.load(os.path.join(os.path.dirname(__file__), "golden", "fit_w2body_plumbing.npz"))
    chem = composition.ChemicalSystem(["W"], degree=2)
    basis = bspline.BSplineBasis(chem, r_min_map={("W", "W"): 0.001}, r_max_map={("W", "W"): 5.5},
                                 resolution_map={("W", "W"): 15}, trailing_trim=3)
    rows = {}
    for k in range(len(fix["energies"])):
        geom = Atoms(numbers=[74] * fix["positions"].shape[1], positions=fix["positions"][k], cell=fix["cell"],
                     pbc=True)
        f = fix["forces"][k]
        rows[f"w_{k}"] = dict(geometry=geom, energy=float(fix["energies"][k]), fx=f[:, 0], fy=f[:, 1], fz=f[:, 2])
    df_features = BasisFeaturizer(basis).evaluate(pd.DataFrame.from_dict(rows, orient="index"), progress=None)
    x_e, y_e, x_f, y_f = ls.dataframe_to_tuples(df_features, n_elements=1, energy_key="energy")
    assert np.allclose(x_e, fix["x_e"], rtol=1e-9, atol=1e-12) and len(y_f) == int(fix["n_force_rows"])
    for solver in ("host", "cusolver"):
        model = ls.WeightedLinearModel(basis, solver=solver, ridge_1b=float(fix["ridge_1b"]),
                                       ridge_2b=float(fix["ridge_2b"]), curvature_2b=float(fix["curvature_2b"]))
        model.fit(x_e, y_e, x_f, y_f, weight=float(fix["weight"]))
        assert np.allclose(model.coefficients, fix["coefficients"], rtol=1e-5, atol=1e-7)
        assert np.allclose(model.predict(x_f), fix["predict_f"], rtol=1e-5, atol=1e-6)


def test_pair_distribution_summary_matches_reference():
    """summarize_distances on the GPU pair list against the histograms the reference wrote for
    the same frames (oracle/make_golden_rdf.py: three triclinic periodic Ne/Xe cells and one
    free cluster, r_cut 7 A, 70 bins)."""
    import os
    from uf3_b200 import composition
    from uf3_b200.distances import summarize_distances
    fix = np.load(os.path.join(os.path.dirname(__file__), "golden", "rdf_summary.npz"))
    frames = []
    for k in range(int(fix["n_frames"])):
        pbc = fix[f"pbc_{k}"]
        frames.append(Atoms(numbers=fix[f"numbers_{k}"], positions=fix[f"positions_{k}"],
                            cell=fix["cell"] if pbc.any() else None, pbc=pbc))
    chem = composition.ChemicalSystem(["Ne", "Xe"], degree=2)
    hist, edges, lower = summarize_distances(frames, chem, r_cut=7.0, n_bins=70, print_stats=False, progress=None)
    assert np.array_equal(edges, fix["edges"])
    for pair in chem.interactions_map[2]:
        key = "-".join(pair)
        assert np.allclose(hist[pair], fix["hist_" + key], rtol=1e-12, atol=0)
        assert lower[pair] == float(fix["lower_" + key])


def test_cusolver_solve_matches_host_lapack():
    rng = np.random.default_rng(3)
    for n in (1, 7, 73, 456):
        m = rng.normal(size=(n + 5, n))
        a = m.T @ m + 1e-3 * np.eye(n)
        b = rng.normal(size=n)
        want = np.linalg.solve(a, b)
        got = ls.device_solve(a, b)
        assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()
    b2 = rng.normal(size=(n, 3))
    assert np.allclose(ls.device_solve(a, b2), np.linalg.solve(a, b2), rtol=1e-8, atol=1e-10)
    a_unsym = rng.normal(size=(9, 9)) + 4 * np.eye(9)      # row-major handling: A x = b, not A^T x = b
    b = rng.normal(size=9)
    assert np.allclose(ls.device_solve(a_unsym, b), np.linalg.solve(a_unsym, b), rtol=1e-10)
    # a model fitted with solver="cusolver" equals the host fit
    case = gu.Case("syn_w54_demo")
    basis = case.basis()
    y_e = float(case["x_energy"] @ np.linspace(-1, 1, basis.n_feats))
    y_f = case["x_forces"] @ np.linspace(-1, 1, basis.n_feats)
    fits = []
    for solver in ("host", "cusolver"):
        model = ls.WeightedLinearModel(basis, solver=solver, ridge_1b=1e-6, ridge_2b=1e-6, ridge_3b=1e-6)
        n = len(case.numbers)
        model.fit(case["x_energy"][None, :] / n, np.array([y_e / n]), case["x_forces"], y_f)
        fits.append(model.coefficients)
    assert np.abs(fits[0] - fits[1]).max() <= 1e-7 * np.abs(fits[0]).max()


def test_frame_pipeline_matches_single_frame_calls():
    from uf3_b200 import pipeline
    names = ["syn_w16_demo", "syn_w54_demo", "syn_w36_slab", "syn_w128_demo", "syn_w16_demo"]
    cases = [gu.Case(n) for n in names]
    feat = BasisFeaturizer(cases[0].basis())
    got = list(pipeline.featurize_frames(feat, [c.atoms() for c in cases]))
    assert len(got) == len(cases)
    for case, (xe, xf) in zip(cases, got):
        assert np.allclose(xe, case["x_energy"], rtol=1e-5, atol=1e-8)
        assert xf.shape == case["x_forces"].shape
        assert np.allclose(xf, case["x_forces"], rtol=1e-5, atol=1e-8)


def test_native_pipeline_matches_single_frame_calls():
    """uf3b_pipeline_*: frames of different sizes in flight on three slots give the rows of the
    one-frame-at-a-time calls, bit for bit; an unknown element surfaces at wait()."""
    import torch
    from uf3_b200 import _native
    from uf3_b200.engine import Engine
    from uf3_b200.pipeline import NativePipeline
    names = ["syn_w54_demo", "syn_w16_demo", "syn_w128_demo", "syn_w36_slab", "syn_w54_demo", "syn_w128_demo", "syn_w16_demo"]
    cases = [gu.Case(n) for n in names]
    basis = cases[0].basis()
    F = basis.n_feats
    pipe = NativePipeline(basis, depth=3)
    eng = Engine(basis)
    pending, outs = [], []
    for case in cases:
        n = len(case.numbers)
        pos = torch.from_numpy(np.ascontiguousarray(case.positions)).pin_memory().numpy()
        num = np.ascontiguousarray(case.numbers, dtype=np.int32)
        xe = torch.empty(F, dtype=torch.float64).pin_memory().numpy()
        xf = torch.empty((3 * n, F), dtype=torch.float64).pin_memory().numpy()
        images = geometry.image_table(case.cell, case.pbc, basis.r_cut)
        pending.append(pipe.submit(pos, num, images, xe, xf))
        outs.append((xe, xf))
        if len(pending) == 3:                 # a ticket is valid until its slot is reused
            pipe.wait(pending.pop(0))
    for t in pending:
        pipe.wait(t)
    with pytest.raises(_native.UF3BError):    # ticket 0 went through slot 0, which has been reused since
        pipe.wait(0)
    for case, (xe, xf) in zip(cases, outs):
        eng.build_neighbors(case.positions, case.numbers, images=geometry.image_table(case.cell, case.pbc, basis.r_cut))
        want_e, want_f = eng.featurize()
        assert np.array_equal(xe, want_e) and np.array_equal(xf, want_f)
        assert gu.rel_err(xf, case["x_forces"]) <= 1e-6
    bad = np.array([74, 26], dtype=np.int32)
    t = pipe.submit(np.zeros((2, 3)), bad, (np.zeros((1, 3), dtype=np.int64), np.zeros((1, 3))), np.zeros(F), np.zeros((6, F)))
    with pytest.raises(_native.ElementError):
        pipe.wait(t)
    pipe.close()
    eng.close()


def test_accumulate_frames_sharded():
    """uf3_b200.distributed.accumulate_frames on one GPU: the two shards add up to the whole."""
    from uf3_b200 import distributed
    cases = [gu.Case(n) for n in ("syn_w16_demo", "syn_w54_demo", "syn_w36_slab")]
    feat = BasisFeaturizer(cases[0].basis())
    rng = np.random.default_rng(2)
    frames = [(c.atoms(), float(rng.normal()), rng.normal(size=(3, len(c.numbers)))) for c in cases]
    whole = distributed.accumulate_frames(feat, frames).to_vector()
    parts = sum(distributed.accumulate_frames(feat, frames, rank=r, world_size=2).to_vector() for r in range(2))
    assert np.allclose(parts, whole, rtol=1e-10, atol=1e-10)
    host = ls.GramStats(feat.engine.n_feats)
    for c, (_, e, f) in zip(cases, frames):
        host.add_energy_row(c["x_energy"], e, len(c.numbers))
        host.add_force_rows(c["x_forces"], f.reshape(-1))
    ref = host.to_vector()
    assert np.abs(whole - ref).max() <= 1e-8 * np.abs(ref).max()


@pytest.mark.gpu
def test_fit_pipeline_matches_host_gram_and_keeps_errors():
    """uf3b_pipeline_submit_fit (distributed.accumulate_frames_pipelined): the rows stay in HBM, the
    normal equations summed over the slots equal the ones built on the host from the reference
    rows; the shards of two ranks add up to the whole; a failing frame nobody waited for is
    reported by a later call instead of being lost."""
    from uf3_b200 import _native, distributed
    from uf3_b200.pipeline import NativePipeline
    names = ("syn_w16_demo", "syn_w54_demo", "syn_w36_slab", "syn_w128_demo", "syn_w54_demo")
    cases = [gu.Case(n) for n in names]
    basis = cases[0].basis()
    F = basis.n_feats
    rng = np.random.default_rng(5)
    frames = [(c.atoms(), float(rng.normal()), rng.normal(size=(3, len(c.numbers)))) for c in cases]
    frames[2] = (frames[2][0], frames[2][1], None)          # an energy-only frame
    whole = distributed.accumulate_frames_pipelined(basis, frames, depth=3).to_vector()
    parts = sum(distributed.accumulate_frames_pipelined(basis, frames, rank=r, world_size=2, depth=2).to_vector()
                for r in range(2))
    assert np.allclose(parts, whole, rtol=1e-10, atol=1e-10)
    host = ls.GramStats(F)
    for c, (_, e, f) in zip(cases, frames):
        host.add_energy_row(c["x_energy"], e, len(c.numbers))
        if f is not None:
            host.add_force_rows(c["x_forces"], f.reshape(-1))
    ref = host.to_vector()
    assert np.abs(whole - ref).max() <= 1e-8 * np.abs(ref).max()
    # sticky error: the bad frame's slot is reused before anybody waits for it
    pipe = NativePipeline(basis, depth=1)
    none = (np.zeros((1, 3), dtype=np.int64), np.zeros((1, 3)))
    pipe.submit_fit(np.zeros((2, 3)), np.array([74, 26], dtype=np.int32), none, None, np.zeros(F))
    with pytest.raises(_native.UF3BError):
        pipe.submit_fit(np.zeros((1, 3)), np.array([74], dtype=np.int32), none, None, np.zeros(F))
    with pytest.raises(_native.UF3BError):
        pipe.export_gram()
    pipe.close()
    # argument validation happens at submit()
    pipe = NativePipeline(basis, depth=2)
    with pytest.raises(_native.UF3BError):
        pipe._native.check(pipe._lib.uf3b_pipeline_submit(pipe._pipe, -1, None, None, 1, None, None, None, None, F, None))
    pipe.close()


@pytest.mark.gpu
def test_singular_normal_equations_are_an_error():
    """uf3b_solve reports a zero pivot of the factorisation instead of returning inf / NaN."""
    from uf3_b200 import _native
    a = np.zeros((4, 4))
    a[0, 0] = a[1, 1] = 1.0
    with pytest.raises(_native.UF3BError):
        ls.device_solve(a, np.ones(4))


def test_deferred_list_builds_keep_the_md_loop_exact():
    """Engine(deferred_lists=True): builds that reuse the cell grid return without a host wait and
    are verified by energy_forces_device; a step that invalidates the cached grid (atoms leave the
    padded box) or overflows the index arrays is repeated transparently.  Energies and forces equal
    those of an engine that checks every build."""
    import torch
    from uf3_b200.engine import Engine
    case = gu.Case("calc_syn_w54_model23")
    basis = case.basis()
    rng = np.random.default_rng(3)
    pos0, numbers, cell, pbc = synthetic.bcc_w((6, 6, 6), a=3.2, sigma=0.05, seed=4)
    images = geometry.image_table(cell, pbc, basis.r_cut)
    fast, safe = Engine(basis, deferred_lists=True), Engine(basis)
    for eng in (fast, safe):
        eng.set_coefficients(case["coefficients"])
    n = len(pos0)
    d_num = torch.from_numpy(numbers).cuda()
    out = {k: (torch.zeros(1, dtype=torch.float64, device="cuda"), torch.zeros((n, 3), dtype=torch.float64, device="cuda"))
           for k in ("fast", "safe")}
    pos = pos0.copy()
    for step in range(8):
        if step == 4:
            pos = pos + np.array([3.0, -2.5, 2.0])          # the whole crystal leaves the cached grid (1 A skin)
        elif step == 6:
            pos = pos * 0.8                                    # denser: more neighbours than the arrays were sized for
            cell = cell * 0.8
            images = geometry.image_table(cell, pbc, basis.r_cut)
        else:
            pos = pos + rng.normal(scale=0.01, size=pos.shape)
        d_pos = torch.from_numpy(np.ascontiguousarray(pos)).cuda()
        for key, eng in (("fast", fast), ("safe", safe)):
            eng.build_neighbors_device(d_pos.data_ptr(), d_num.data_ptr(), n, images)
            eng.energy_forces_device(out[key][0].data_ptr(), out[key][1].data_ptr())
        torch.cuda.synchronize()
        e_f, e_s = float(out["fast"][0]), float(out["safe"][0])
        assert abs(e_f - e_s) <= 1e-12 * abs(e_s), step
        assert gu.rel_err(out["fast"][1].cpu().numpy(), out["safe"][1].cpu().numpy()) <= 1e-11, step
    assert fast.neighbor_count(3) == safe.neighbor_count(3)
    fast.close()
    safe.close()


def test_pipeline_repeats_frames_whose_deferred_lists_were_invalid():
    """Same-size frames reuse their slot's cell grid and run without host waits; a frame whose 3-body
    rows outgrow the previous frame's longest one, or whose atoms leave the cached grid, is detected
    by the device-side flag and repeated — rows and normal equations equal the checked calls."""
    import torch
    from uf3_b200.engine import Engine
    from uf3_b200.pipeline import NativePipeline
    basis = synthetic.w_basis("demo")
    F = basis.n_feats
    specs = [(3.165, 0.05, 0.0), (3.165, 0.05, 0.0), (3.165, 0.05, 0.0), (2.95, 0.05, 0.0), (3.165, 0.05, 0.0),
             (3.165, 0.05, 4.0), (3.165, 0.2, 4.0), (3.165, 0.05, 0.0)]
    frames = []
    for k, (a, sigma, shift) in enumerate(specs):
        pos, numbers, cell, pbc = synthetic.bcc_w((4, 4, 5), a=a, sigma=sigma, seed=20 + k)
        frames.append((np.ascontiguousarray(pos + shift), numbers, cell, pbc))
    rng = np.random.default_rng(9)
    eng = Engine(basis)
    for depth in (1, 2):
        pipe = NativePipeline(basis, depth=depth)
        outs, pending = [], []
        for pos, numbers, cell, pbc in frames:
            n = len(pos)
            xe = torch.empty(F, dtype=torch.float64).pin_memory().numpy()
            xf = torch.empty((3 * n, F), dtype=torch.float64).pin_memory().numpy()
            pending.append(pipe.submit(pos, numbers, geometry.image_table(cell, pbc, basis.r_cut), xe, xf))
            outs.append((xe, xf))
            if len(pending) == depth:
                pipe.wait(pending.pop(0))
        for t in pending:
            pipe.wait(t)
        host = ls.GramStats(F)
        ys = []
        for (pos, numbers, cell, pbc), (xe, xf) in zip(frames, outs):
            eng.build_neighbors(pos, numbers, images=geometry.image_table(cell, pbc, basis.r_cut))
            want_e, want_f = eng.featurize()
            assert np.array_equal(xe, want_e) and np.array_equal(xf, want_f)
            y = rng.normal(size=3 * len(pos))
            ys.append(y)
            host.add_force_rows(want_f, y)
        # fit mode over the same frames
        pending = []
        for (pos, numbers, cell, pbc), y in zip(frames, ys):
            pending.append(pipe.submit_fit(pos, numbers, geometry.image_table(cell, pbc, basis.r_cut), y, np.zeros(F)))
            if len(pending) == depth:
                pipe.wait(pending.pop(0))
        gram, ordinate, moments = pipe.export_gram()
        assert np.abs(gram - host.gram_f).max() <= 1e-9 * np.abs(host.gram_f).max()
        assert np.abs(ordinate - host.ord_f).max() <= 1e-9 * np.abs(host.ord_f).max()
        assert moments[0] == sum(len(y) for y in ys)
        pipe.close()
    eng.close()


def test_deferred_feature_rows_are_verified_at_the_next_build():
    """Engine(deferred_lists=True) with device outputs (the resident arm of bench.py): the row kernels run
    behind an unverified list build; the slot's NEXT build verifies it and raises (UF3B_RETRY) if the rows
    that were produced from it are invalid, and neighbor_count does the same for the last frame."""
    import torch
    from uf3_b200 import _native
    from uf3_b200.engine import Engine
    basis = synthetic.w_basis("demo")
    F = basis.n_feats
    pos, numbers, cell, pbc = synthetic.bcc_w((4, 4, 5), seed=31)
    images = geometry.image_table(cell, pbc, basis.r_cut)
    n = len(pos)
    safe = Engine(basis)
    safe.build_neighbors(pos, numbers, images=images)
    want_e, want_f = safe.featurize()
    eng = Engine(basis, deferred_lists=True)
    d_num = torch.from_numpy(numbers).cuda()
    xe = torch.empty(F, dtype=torch.float64, device="cuda")
    xf = torch.empty((3 * n, F), dtype=torch.float64, device="cuda")

    def frame(p):
        d_pos = torch.from_numpy(np.ascontiguousarray(p)).cuda()
        eng.build_neighbors_device(d_pos.data_ptr(), d_num.data_ptr(), n, images)
        eng.featurize_device(xe.data_ptr(), xf.data_ptr(), F)
        torch.cuda.synchronize()

    frame(pos)                                   # first build of the handle: checked
    frame(pos + 0.01)                            # deferred, valid
    assert np.allclose(xf.cpu().numpy(), want_f, rtol=1e-6, atol=1e-6 * np.abs(want_f).max())
    frame(pos)                                   # verifies the previous build: fine
    assert np.array_equal(xf.cpu().numpy(), want_f) and np.array_equal(xe.cpu().numpy(), want_e)
    frame(pos + 4.0)                             # the atoms leave the cached grid: rows of this frame are invalid
    with pytest.raises(_native.UF3BError) as err:
        frame(pos)                               # ... and the next build says so
    assert err.value.code == _native.RETRY
    frame(pos)                                   # the grid was dropped: a checked build
    assert np.array_equal(xf.cpu().numpy(), want_f)
    frame(pos + 4.0)
    with pytest.raises(_native.UF3BError):
        eng.neighbor_count(3)                    # the last frame of a stream is verified the same way
    eng.close()
    safe.close()
