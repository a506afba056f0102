"""Pair-distribution analysis on the GPU neighbour list.

`summarize_distances` keeps the signature and the return values of
`/root/reference/uf3/representation/distances.py:367-442` (histogram of pair distances per
pair interaction over a list of geometries, divided by 4 pi r^2, density and entry count;
lower bound = left edge of the first populated bin).  The reference masks a dense
(atoms x supercell) distance matrix per interaction and calls `np.histogram`; here Kernel A
builds the pair list once with bounds (0, r_cut) and `uf3b_pair_histogram` counts every
(centre, neighbour) entry into its (pair, bin) counter.
"""
import ctypes as C

import numpy as np

from uf3_b200 import _native, bspline, composition, geometry
from uf3_b200.atoms import frame_arrays


def _counting_engine(chemical_system, r_cut, device=None):
    from uf3_b200.engine import Engine
    pairs = chemical_system.interactions_map[2]
    chem2 = composition.ChemicalSystem(list(chemical_system.element_list), degree=2)
    basis = bspline.BSplineBasis(chem2, r_min_map={pair: 0.0 for pair in pairs},
                                 r_max_map={pair: float(r_cut) for pair in pairs},
                                 resolution_map={pair: 4 for pair in pairs})
    return Engine(basis, device=device), basis


def pair_histogram_counts(engine, bin_edges):
    """int64 counts [n_pairs, n_bins] of the list-2 entries of the engine's current frame."""
    edges = np.ascontiguousarray(bin_edges, dtype=np.float64)
    n_bins = len(edges) - 1
    ne = len(engine.tables.element_list)
    n_pairs = ne * (ne + 1) // 2
    counts = np.zeros((n_pairs, n_bins), dtype=np.int64)
    _native.check(engine._lib.uf3b_pair_histogram(engine._basis, engine._nlist, C.c_void_p(edges.ctypes.data),
                                                  n_bins, C.c_void_p(counts.ctypes.data), None))
    return counts


def summarize_distances(geometries, chemical_system, r_cut=12.0, n_bins=100, print_stats=True,
                        min_peak_width=0.5, progress="bar", device=None):
    pair_tuples = chemical_system.interactions_map[2]
    bin_edges = np.linspace(0, r_cut, n_bins + 1)
    histogram_values = {pair: np.zeros(n_bins) for pair in pair_tuples}
    n_entries = len(geometries)
    engine, basis = _counting_engine(chemical_system, r_cut, device)
    order = list(basis.interactions_map[2])          # the device's pair order
    try:
        for geom in geometries:
            positions, numbers, cell, pbc = frame_arrays(geom)
            if np.any(pbc):
                images = geometry.image_table(cell, pbc, r_cut)
                density = len(positions) / abs(np.linalg.det(np.asarray(cell, dtype=np.float64)))
            else:
                images = None
                density = 1
            engine.build_neighbors(positions, numbers, images=images)
            counts = pair_histogram_counts(engine, bin_edges)
            for pair in pair_tuples:
                frequencies = counts[order.index(pair)] / density / n_entries / 2
                if pair[0] != pair[1]:
                    frequencies = frequencies / 2
                histogram_values[pair] += frequencies
    finally:
        engine.close()
    bin_centers = 0.5 * np.add(bin_edges[:-1], bin_edges[1:])
    bin_span = int(np.ceil(min_peak_width / (bin_edges[1] - bin_edges[0])))
    lower_bounds = {}
    for pair in pair_tuples:
        histogram_values[pair] /= bin_centers ** 2 * 4 * np.pi
        lower_bound = bin_edges[np.nonzero(histogram_values[pair])[0][0]]
        lower_bounds[pair] = lower_bound
        if print_stats:
            from scipy import signal
            peaks = bin_centers[signal.find_peaks(histogram_values[pair], width=bin_span)[0]]
            print(pair, "Lower bound: {0:.3f} angstroms".format(lower_bound))
            print(pair, "Peaks (min width {} angstroms):".format(min_peak_width), peaks)
    return histogram_values, bin_edges, lower_bounds
