#!/usr/bin/env python
"""Golden pair-distribution histograms written by the UNMODIFIED reference
(distances.summarize_distances, representation/distances.py:367-442; build container only):

    python oracle/make_golden_rdf.py        # writes tests/golden/rdf_summary.npz

TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(HERE, "ref_shims"), "/root/reference"]
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
warnings.simplefilter("ignore")

import ase  # noqa: E402 (stand-in)
from uf3.data import composition  # noqa: E402
from uf3.representation import distances  # noqa: E402


def main():
    rng = np.random.default_rng(21)
    cell = np.array([[6.1, 0.4, 0.0], [0.9, 5.7, 0.3], [-0.5, 0.8, 6.4]])
    frames = []
    for k in range(3):
        frac = rng.random((14, 3))
        frames.append(ase.Atoms(numbers=[10] * 8 + [54] * 6, positions=frac @ cell, cell=cell, pbc=True))
    frames.append(ase.Atoms(numbers=[10, 54, 10, 54], positions=rng.random((4, 3)) * 5.0, pbc=False))
    chem = composition.ChemicalSystem(["Ne", "Xe"], degree=2)
    hist, edges, lower = distances.summarize_distances(frames, chem, r_cut=7.0, n_bins=70, print_stats=False,
                                                       progress=None)
    rec = dict(edges=edges, cell=cell, n_frames=len(frames))
    for k, fr in enumerate(frames):
        rec[f"positions_{k}"] = fr.get_positions()
        rec[f"numbers_{k}"] = fr.get_atomic_numbers()
        rec[f"pbc_{k}"] = np.array(fr.get_pbc())
    for pair, values in hist.items():
        rec["hist_" + "-".join(pair)] = values
        rec["lower_" + "-".join(pair)] = lower[pair]
    out = os.path.join(REPO, "tests", "golden", "rdf_summary.npz")
    np.savez_compressed(out, **rec)
    print("wrote", out, {k: float(v) for k, v in lower.items()})


if __name__ == "__main__":
    main()
