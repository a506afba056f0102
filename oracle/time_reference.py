#!/usr/bin/env python
"""Time the UNMODIFIED reference featurizer on the benchmark's own frame generator (build
container only; the reference cannot travel to the GPU box and cannot hold frames above
~2 000 atoms: dense M x M distance matrices).

    python oracle/time_reference.py [--parallel] > profiles/r02_reference_as_is_cpu.json

Rattled bcc W (a = 3.165 A, sigma = 0.05 A), demo 2+3-body basis (73 columns), energy row +
3N force rows through `BasisFeaturizer.evaluate_configuration` (process.py:293), one process,
one core, numba warmed; median of the repeats.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.
"""
import json
import os
import statistics
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "ref_shims"), "/root/reference"]
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
warnings.simplefilter("ignore")

import ase  # noqa: E402 (stand-in)
from uf3.data import composition  # noqa: E402
from uf3.representation import bspline, process  # noqa: E402


def bcc_w(reps, a=3.165, sigma=0.05, seed=0):
    rng = np.random.default_rng(seed)
    base = np.array([[0, 0, 0], [0.5, 0.5, 0.5]]) * a
    cells = np.array([[i, j, k] for i in range(reps[0]) for j in range(reps[1]) for k in range(reps[2])]) * a
    pos = (cells[:, None, :] + base[None, :, :]).reshape(-1, 3) + rng.normal(0, sigma, (2 * len(cells), 3))
    return ase.Atoms(numbers=[74] * len(pos), positions=pos, cell=np.diag(np.array(reps) * a), pbc=True)


def main():
    chem = composition.ChemicalSystem(["W"], degree=3)
    basis = bspline.BSplineBasis(
        chem, r_min_map={("W", "W"): 0.001, ("W", "W", "W"): [1.5, 1.5, 1.5]},
        r_max_map={("W", "W"): 5.5, ("W", "W", "W"): [3.5, 3.5, 7.0]},
        resolution_map={("W", "W"): 15, ("W", "W", "W"): [6, 6, 12]},
        leading_trim={2: 0, 3: 3}, trailing_trim={2: 3, 3: 3})
    feat = process.BasisFeaturizer(basis)
    out = {"what": "reference as-is (uf3 @ /root/reference), BasisFeaturizer.evaluate_configuration, "
                   "energy row + 3N force rows, demo 2+3-body basis, 1 process / 1 core",
           "host": {"cpus": os.cpu_count()}, "n_feats": int(basis.n_feats), "runs": []}
    feat.evaluate_configuration(bcc_w((2, 2, 2)), energy=0.0, forces=np.zeros((3, 16)))   # numba warm-up
    for reps, repeats in (((3, 3, 3), 5), ((4, 4, 4), 5), ((5, 5, 5), 3), ((6, 6, 6), 1)):
        geom = bcc_w(reps, seed=1)
        n = len(geom)
        times = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            feat.evaluate_configuration(geom, energy=0.0, forces=np.zeros((3, n)))
            times.append(time.perf_counter() - t0)
        t = statistics.median(times)
        out["runs"].append({"n_atoms": n, "seconds_per_frame": t, "atom_steps_per_s": n / t, "repeats": repeats})
        print(f"{n} atoms: {t:.3f} s/frame, {n / t:.1f} atom-steps/s", file=sys.stderr)
    if "--parallel" in sys.argv:
        out["parallel"] = parallel_runs(feat)
    print(json.dumps(out, indent=1))


def parallel_runs(feat):
    """The reference's own parallel mode: `BasisFeaturizer.evaluate_parallel`
    (process.py:196-254) over a `ProcessPoolExecutor(max_workers=cores)`, 4 x cores frames of 128 atoms
    (the size its example data set has), every host core busy."""
    import pandas as pd
    from concurrent.futures import ProcessPoolExecutor
    cores = os.cpu_count() or 1
    n_frames = 4 * cores
    rows = {}
    for k in range(n_frames):
        geom = bcc_w((4, 4, 4), seed=100 + k)
        n = len(geom)
        rows[f"w_{k}"] = {"geometry": geom, "energy": 0.0, "fx": np.zeros(n), "fy": np.zeros(n), "fz": np.zeros(n)}
    df = pd.DataFrame.from_dict(rows, orient="index")
    runs = []
    with ProcessPoolExecutor(max_workers=cores) as pool:
        feat.evaluate_parallel(df.iloc[:cores], pool, n_jobs=cores, progress=False)        # numba warm-up per worker
        t0 = time.perf_counter()
        table = feat.evaluate_parallel(df, pool, n_jobs=cores, progress=False)
        t = time.perf_counter() - t0
    assert len(table) == n_frames * (1 + 3 * n)
    runs.append({"n_atoms": n, "frames": n_frames, "workers": cores, "seconds": t,
                 "atom_steps_per_s": n_frames * n / t})
    print(f"parallel: {n_frames} frames x {n} atoms on {cores} workers: {t:.1f} s, {n_frames * n / t:.1f} atom-steps/s",
          file=sys.stderr)
    return runs


if __name__ == "__main__":
    main()
