#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference (build container only).

    python oracle/make_golden.py            # writes tests/golden/*.npz

Imports the reference from /root/reference through the stand-in third-party
modules in oracle/ref_shims (ase / ndsplines / tables are not installed here),
runs its own featurizer (`BasisFeaturizer.evaluate_configuration`,
process.py:293) and evaluator (`UFCalculator`, calculator.py:40) on
 (a) the geometries and basis settings of the reference's own hot-path tests
     (tests/test_representation.py, tests/test_calculator.py,
     tests/test_distances.py, tests/test_optimize.py), and
 (b) seeded synthetic frames (rattled bcc W, Ne/Xe, a triclinic ternary cell),
and stores inputs + outputs as small .npz fixtures.  The GPU box has no
/root/reference: tests there read only the committed fixtures.

TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(HERE, "ref_shims"), "/root/reference"]
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
warnings.simplefilter("ignore")

import ase  # noqa: E402  (stand-in)
from uf3.data import composition, geometry  # noqa: E402
from uf3.representation import bspline, process, distances, angles  # noqa: E402
from uf3.regression import least_squares  # noqa: E402
from uf3.forcefield import calculator  # noqa: E402
from uf3.util import json_io  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")
REF_DATA = "/root/reference/tests/data/precalculated_ref"


def enc(obj):
    """kwargs with tuple keys -> JSON-able (dash-joined keys, lists)."""
    if isinstance(obj, dict):
        return {("-".join(map(str, k)) if isinstance(k, tuple) else str(k)): enc(v)
                for k, v in obj.items()}
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    if isinstance(obj, (list, tuple)):
        return [enc(v) for v in obj]
    if isinstance(obj, np.generic):
        return obj.item()
    return obj


def reference_pair_list(geom, basis):
    """np.where of the union of the reference's per-interaction masks, real centres
    (distances.py:48-69)."""
    if any(geom.pbc):
        sup = geometry.get_supercell(geom, r_cut=basis.r_cut)
    else:
        sup = geom
    dm = distances.get_distance_matrix(geom, sup)
    gz = np.array(geom.get_atomic_numbers())
    sz = np.array(sup.get_atomic_numbers())
    mask = np.zeros_like(dm, dtype=bool)
    for pair in basis.interactions_map[2]:
        nums = ase.symbols.symbols2numbers(pair)
        comp = distances.mask_matrix_by_pair_interaction(nums, gz, sz)
        r_min = max(basis.r_min_map[pair], 0)
        mask |= comp & (dm > r_min) & (dm < basis.r_max_map[pair])
    i, j = np.where(mask)
    return i.astype(np.int64), j.astype(np.int64), len(sup)


def reference_trio_list(geom, basis):
    """identify_ij(square=False) (angles.py:289-342)."""
    if basis.degree < 3:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    sup = geometry.get_supercell(geom, r_cut=basis.r_cut) if any(geom.pbc) else geom
    knot_sets = [basis.knots_map[t] for t in basis.interactions_map[3]]
    _, i, j = angles.identify_ij(geom, knot_sets, sup)
    return i.astype(np.int64), j.astype(np.int64)


def featurize_case(name, geom, els, degree, kwargs, forces=True, lists=True):
    chem = composition.ChemicalSystem(els, degree=degree)
    basis = bspline.BSplineBasis(chem, **kwargs)
    feat = process.BasisFeaturizer(basis)
    n = len(geom)
    rows = feat.evaluate_configuration(
        geom, energy=0.0, forces=np.zeros((3, n)) if forces else None)
    rec = dict(kind="featurize", name=name,
               positions=geom.get_positions(), numbers=geom.get_atomic_numbers(),
               cell=np.array(geom.get_cell()), pbc=np.array(geom.get_pbc()),
               config=json.dumps(dict(element_list=list(els), degree=degree,
                                      kwargs=enc(kwargs))),
               columns=json.dumps(basis.get_column_names()),
               x_energy=rows["energy"][1:])
    if forces:
        xf = np.stack([rows[f"{c}_{a}"][1:] for c in ("fx", "fy", "fz")
                       for a in range(n)])
        rec["x_forces"] = xf
    if lists:
        i2, j2, n_sup = reference_pair_list(geom, basis)
        i3, j3 = reference_trio_list(geom, basis)
        rec.update(nl2_i=i2, nl2_j=j2, nl3_i=i3, nl3_j=j3, n_sup=n_sup)
    save(name, rec)
    return basis


def calculator_case(name, geom, model, lists=False):
    calc = calculator.UFCalculator(model)
    geom.calc = calc
    e = geom.get_potential_energy()
    f = geom.get_forces()
    basis = model.bspline_config
    rec = dict(kind="calculator", name=name,
               positions=geom.get_positions(), numbers=geom.get_atomic_numbers(),
               cell=np.array(geom.get_cell()), pbc=np.array(geom.get_pbc()),
               config=json.dumps(dict(
                   element_list=list(basis.element_list), degree=basis.degree,
                   kwargs=enc(dict(knots_map=basis.knots_map,
                                   knot_strategy=basis.knot_strategy,
                                   leading_trim=basis.leading_trim,
                                   trailing_trim=basis.trailing_trim)))),
               coefficients=np.asarray(model.coefficients, dtype=float),
               energy=float(e), forces=np.asarray(f))
    save(name, rec)


def save(name, rec):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **rec)
    print(f"{name:28s} {os.path.getsize(path) / 1024:8.1f} KiB")


# --------------------------------------------------------------- geometries
def bcc_w(reps, a=3.165, sigma=0.05, seed=0):
    rng = np.random.default_rng(seed)
    base = np.array([[0, 0, 0], [0.5, 0.5, 0.5]])
    cells = np.array([[i, j, k] for i in range(reps[0]) for j in range(reps[1])
                      for k in range(reps[2])])
    pos = (cells[:, None, :] + base[None, :, :]).reshape(-1, 3) * a
    pos = pos + rng.normal(0, sigma, pos.shape)
    return ase.Atoms(numbers=[74] * len(pos), positions=pos,
                     cell=np.diag(np.array(reps) * a), pbc=True)


def nexe(reps, a=8.0, sigma=0.1, seed=0):
    """fcc Ne at the origin + fcc Xe shifted by a/2 along x (generate_Ne-Xe.in:6-12)."""
    rng = np.random.default_rng(seed)
    fcc = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    cells = np.array([[i, j, k] for i in range(reps[0]) for j in range(reps[1])
                      for k in range(reps[2])])
    ne = (cells[:, None, :] + fcc[None]).reshape(-1, 3)
    xe = ne + np.array([0.5, 0, 0])
    pos = np.concatenate([ne, xe]) * a
    pos = pos + rng.normal(0, sigma, pos.shape)
    numbers = [10] * len(ne) + [54] * len(xe)
    return ase.Atoms(numbers=numbers, positions=pos,
                     cell=np.diag(np.array(reps) * a), pbc=True)


W_DEMO = dict(r_min_map={("W", "W"): 0.001, ("W", "W", "W"): [1.5, 1.5, 1.5]},
              r_max_map={("W", "W"): 5.5, ("W", "W", "W"): [3.5, 3.5, 7.0]},
              resolution_map={("W", "W"): 15, ("W", "W", "W"): [6, 6, 12]},
              leading_trim={2: 0, 3: 3}, trailing_trim={2: 3, 3: 3})
W_MANUSCRIPT = dict(r_min_map={("W", "W"): 1.5, ("W", "W", "W"): [1.5, 1.5, 1.5]},
                    r_max_map={("W", "W"): 5.5, ("W", "W", "W"): [4.25, 4.25, 8.5]},
                    resolution_map={("W", "W"): 22, ("W", "W", "W"): [10, 10, 20]},
                    leading_trim={2: 0, 3: 3}, trailing_trim={2: 3, 3: 3})


def main():
    # ---- (a) the reference's own test geometries --------------------------------
    h2o = ase.Atoms("H2O", positions=[[0, 0, 0], [1.5, 0, 0], [0, 2.0, 0]], pbc=False)
    ch4 = ase.Atoms("CH4", positions=[[15.0, 15.0, 15.000010729],
                                      [15.629117489, 15.629117489, 15.629128218],
                                      [14.370881617, 14.370881617, 15.629128218],
                                      [15.629117489, 14.370881617, 14.370892346],
                                      [14.370881617, 15.629117489, 14.370892346]],
                    pbc=True, cell=[30, 30, 30])
    steel = ase.Atoms("Fe8C3", positions=[
        [1.99342831e-01, 7.23471398e-02, 2.29537708e-01],
        [3.27460597e+00, 3.16932506e-03, -9.68273914e-02],
        [3.65842563e-01, 3.07348695e+00, -1.43894877e-01],
        [3.02851201e+00, 2.85731646e+00, 6.85404929e-03],
        [-1.60754569e-03, -3.82656049e-01, 2.57501643e+00],
        [2.80754249e+00, -3.02566224e-01, 2.88284947e+00],
        [-8.16048151e-02, 2.53753926e+00, 3.26312975e+00],
        [2.92484474e+00, 2.93350564e+00, 2.58505036e+00],
        [1.32612346e+00, 1.45718452e+00, -1.80198715e-01],
        [1.51013960e+00, -7.01277380e-02, 1.37666125e+00],
        [-7.03413224e-02, 1.80545564e+00, 1.43230056e+00]],
        pbc=True, cell=[5.74, 5.74, 5.74])
    for tag, lead, trail in (("trimA", {2: 0, 3: 3}, {2: 3, 3: 3}),
                             ("trimB", {2: 0, 3: 0}, {2: 3, 3: 3})):
        featurize_case(f"ref_h2o_{tag}", h2o, ["H", "O"], 3,
                       dict(leading_trim=lead, trailing_trim=trail))
        featurize_case(f"ref_ch4_{tag}", ch4, ["H", "C"], 3,
                       dict(leading_trim=lead, trailing_trim=trail))
    trios = [("Fe", "Fe", "Fe"), ("Fe", "Fe", "C"), ("Fe", "C", "C"),
             ("C", "Fe", "Fe"), ("C", "Fe", "C"), ("C", "C", "C")]
    pairs = [("Fe", "Fe"), ("Fe", "C"), ("C", "C")]
    steel_kwargs = dict(
        r_min_map={**{p: 0.1 for p in pairs}, **{t: [1.5, 1.5, 1.5] for t in trios}},
        r_max_map={**{p: 6.0 for p in pairs}, **{t: [5.0, 5.0, 10.0] for t in trios}},
        resolution_map={**{p: 12 for p in pairs}, **{t: [4, 4, 8] for t in trios}},
        knot_strategy="linear", offset_1b=True, leading_trim=0, trailing_trim=3)
    featurize_case("ref_steel_pbc", steel, ["Fe", "C"], 3, steel_kwargs)
    # cross-check against the reference's committed golden rows (max rel. error 0)
    with open(os.path.join(REF_DATA, "rattled_steel_features.json")) as fh:
        ref_rows = json.load(fh)
    got = np.load(os.path.join(OUT, "ref_steel_pbc.npz"))
    assert np.allclose(got["x_energy"], np.array(ref_rows["energy"])[1:], rtol=1e-12, atol=0)
    tri = [[2, 0, 0], [3, 1.5, 0], [0.5, 0, 2.5]]
    au2 = ase.Atoms("Au2", positions=[[0, 0, 0], [0.5, 0.3, 0.2]], pbc=True, cell=tri)
    nx2 = ase.Atoms("NeXe", positions=[[0, 0, 0], [0.5, 0.3, 0.2]], pbc=True, cell=tri)
    featurize_case("ref_au2_triclinic", au2, ["Au"], 2,
                   dict(r_min_map={("Au", "Au"): 0.5}, r_max_map={("Au", "Au"): 3.0}))
    featurize_case("ref_nexe_triclinic", nx2, ["Ne", "Xe"], 2,
                   dict(r_min_map={("Ne", "Ne"): 0.5, ("Ne", "Xe"): 0.6, ("Xe", "Xe"): 0.7},
                        r_max_map={("Ne", "Ne"): 3.0, ("Ne", "Xe"): 4.0, ("Xe", "Xe"): 3.5}))
    ar3 = ase.Atoms("Ar3", positions=[[0, 0, 0], [3, 0, 0], [0, 4, 0]], pbc=False)
    featurize_case("ref_ar3_default", ar3, ["Ar"], 2, {})

    # ---- evaluator known-answer cases (tests/test_calculator.py) ----------------
    chem = composition.ChemicalSystem(["W"])
    basis = bspline.BSplineBasis(chem, r_min_map={("W", "W"): 2.0},
                                 r_max_map={("W", "W"): 6.0},
                                 resolution_map={("W", "W"): 20}, knot_strategy="lammps")
    model = least_squares.WeightedLinearModel(bspline_config=basis)
    x = np.linspace(2.0, 6.0, 1000)
    y = 4 * 0.87 * ((2.5 / x) ** 12 - (2.5 / x) ** 6)
    model.coefficients = np.insert(
        bspline.fit_spline_1d(x, y, basis.knots_map[("W", "W")]), 0, 0)
    dimer = ase.Atoms("W2", positions=[[0, 0, 0], [1.5, 1.5, 1.5]], pbc=False)
    calculator_case("calc_w_dimer_free", dimer, model)
    dimer_pbc = ase.Atoms("W2", positions=[[0, 0, 0], [1.5, 1.5, 1.5]], pbc=True,
                          cell=[[3, 0, 0], [3, 5, 0], [0, 0, 3]])
    calculator_case("calc_w_dimer_pbc", dimer_pbc, model)
    unary = least_squares.WeightedLinearModel.from_json(
        os.path.join(REF_DATA, "model_unary.json"))
    trimer = ase.Atoms("W3", positions=[[0, 0, 0], [2, 0, 0], [0, 3, 0]], pbc=False)
    calculator_case("calc_w_trimer", trimer, unary)
    w8 = ase.Atoms("W8", positions=[[0.00, 0.00, 0.00], [2.89, 0.12, -0.04],
                                    [-0.32, 2.71, -0.11], [2.65, 2.81, 0.37],
                                    [0.00, 0.00, 3.00], [2.64, 0.00, 3.00],
                                    [-0.08, 2.94, 3.16], [2.53, 2.87, 3.23]],
                   pbc=True, cell=np.eye(3) * 2.74 * 2)
    calculator_case("calc_w8_pbc", w8, unary)
    binary = least_squares.WeightedLinearModel.from_json(
        os.path.join(REF_DATA, "model_binary.json"))
    nx = ase.Atoms("NeXe", positions=[[0, 0, 0], [3.1, 0, 0]], pbc=False)
    calculator_case("calc_nexe_dimer", nx, binary)

    # ---- (b) seeded synthetic frames --------------------------------------------
    featurize_case("syn_w16_demo", bcc_w((2, 2, 2), seed=1), ["W"], 3, W_DEMO)
    featurize_case("syn_w54_demo", bcc_w((3, 3, 3), seed=2), ["W"], 3, W_DEMO)
    featurize_case("syn_w128_demo", bcc_w((4, 4, 4), seed=3), ["W"], 3, W_DEMO)
    featurize_case("syn_w54_manuscript", bcc_w((3, 3, 3), seed=4), ["W"], 3, W_MANUSCRIPT)
    featurize_case("syn_w432_demo_energy", bcc_w((6, 6, 6), seed=5), ["W"], 3, W_DEMO,
                   forces=False)
    featurize_case("syn_w36_slab", bcc_w((3, 3, 2), seed=6), ["W"], 3, W_DEMO)
    nexe_kwargs = dict(
        r_min_map={("Ne", "Ne"): 2.0, ("Ne", "Xe"): 2.5, ("Xe", "Xe"): 3.0},
        r_max_map={("Ne", "Ne"): 6.0, ("Ne", "Xe"): 7.0, ("Xe", "Xe"): 8.0},
        resolution_map={("Ne", "Ne"): 15, ("Ne", "Xe"): 15, ("Xe", "Xe"): 15})
    featurize_case("syn_nexe64_pair", nexe((2, 2, 2), seed=7), ["Ne", "Xe"], 2, nexe_kwargs)
    rng = np.random.default_rng(11)
    cellt = np.array([[6.1, 0.4, 0.0], [0.9, 5.7, 0.3], [-0.5, 0.8, 6.4]])
    frac = rng.random((18, 3))
    tern = ase.Atoms(numbers=[1] * 7 + [6] * 6 + [8] * 5, positions=frac @ cellt,
                     cell=cellt, pbc=True)
    featurize_case("syn_ternary_triclinic", tern, ["H", "C", "O"], 3,
                   dict(r_min_map={("H", "H"): 0.3, ("C", "H"): 0.4},
                        r_max_map={("H", "H"): 4.5, ("C", "O"): 5.0,
                                   ("C", "H", "O"): [3.0, 3.6, 5.5]},
                        resolution_map={("C", "H", "O"): [4, 5, 7]},
                        leading_trim={2: 1, 3: 0}, trailing_trim={2: 2, 3: 1}))
    mixed = ase.Atoms(numbers=[1] * 7 + [6] * 6 + [8] * 5, positions=frac @ cellt,
                      cell=cellt, pbc=[True, False, True])
    featurize_case("syn_ternary_mixed_pbc", mixed, ["H", "C", "O"], 3, {})

    # evaluator on synthetic frames with the shipped example models
    w23 = least_squares.WeightedLinearModel.from_json(
        "/root/reference/examples/tungsten_extxyz/model_2and3.json")
    calculator_case("calc_syn_w54_model23", bcc_w((3, 3, 3), seed=8), w23)
    calculator_case("calc_syn_w128_model23", bcc_w((4, 4, 4), seed=9, sigma=0.15), w23)
    pairm = least_squares.WeightedLinearModel.from_json(
        "/root/reference/examples/NeXe_lammps/model_pair.json")
    calculator_case("calc_syn_nexe64_pair", nexe((2, 2, 2), seed=10), pairm)


if __name__ == "__main__":
    main()
