#!/usr/bin/env python
"""Fixture for the one documented deviation from the reference: a trio whose two
neighbours have the SAME species but different l / m knot vectors (symmetry 1 by
bspline.py:723-763).  The reference decides which neighbour is leg l by supercell index
(angles.py:460-488); for ghost-centred triangles of its force path that order need not agree with
the order the same triangle has when seen from its real centre (the energy path), so its force rows
stop being the derivative of its own energy row.

    python oracle/make_golden_sym1.py        # writes tests/golden/dev_w16_sym1.npz

Stored: the reference's energy row and force rows at the frame, and its ENERGY ROW at positions
displaced by +-delta for a few (atom, direction) pairs, so that tests can take central differences of
the reference's own energy features without the reference.  TEST INFRASTRUCTURE ONLY (build container).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (sets up the stand-in modules and the reference import path)
from make_golden import ase, bspline, composition, process  # noqa: E402

KWARGS = dict(r_min_map={("W", "W"): 0.001, ("W", "W", "W"): [1.5, 1.5, 1.5]},
              r_max_map={("W", "W"): 5.5, ("W", "W", "W"): [3.5, 4.2, 7.0]},
              resolution_map={("W", "W"): 15, ("W", "W", "W"): [6, 8, 12]},
              leading_trim={2: 0, 3: 3}, trailing_trim={2: 3, 3: 3})
DELTA = 1e-5
PROBES = [(0, 0), (3, 1), (7, 2), (12, 0)]          # (atom, direction)


def main():
    geom = mg.bcc_w((2, 2, 2), seed=21, sigma=0.08)
    basis = mg.featurize_case("dev_w16_sym1", geom, ["W"], 3, KWARGS)
    assert basis.symmetry[("W", "W", "W")] == 1
    feat = process.BasisFeaturizer(basis)
    plus, minus = [], []
    for atom, axis in PROBES:
        for sign, out in ((1.0, plus), (-1.0, minus)):
            moved = geom.copy()
            pos = moved.get_positions()
            pos[atom, axis] += sign * DELTA
            moved.set_positions(pos)
            out.append(feat.evaluate_configuration(moved, energy=0.0)["energy"][1:])
    path = os.path.join(mg.OUT, "dev_w16_sym1.npz")
    rec = dict(np.load(path))
    rec.update(probe_atoms=np.array([p[0] for p in PROBES]), probe_axes=np.array([p[1] for p in PROBES]),
               probe_delta=DELTA, x_energy_plus=np.array(plus), x_energy_minus=np.array(minus))
    np.savez_compressed(path, **rec)
    n = len(geom)
    fd = -(np.array(plus) - np.array(minus)) / (2 * DELTA)
    rows = np.stack([rec["x_forces"][axis * n + atom] for atom, axis in PROBES])
    err = np.abs(rows - fd).max(axis=0)
    print("columns where the reference's force rows differ from -d(energy row)/dR:",
          int((err > 1e-4 * np.abs(fd).max()).sum()), "of", fd.shape[1], " max abs diff", err.max())
    print(json.dumps({"n_feats": int(fd.shape[1])}))


if __name__ == "__main__":
    main()
