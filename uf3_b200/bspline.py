"""Cubic B-spline basis descriptor (host side) — the API kept from the reference.

`BSplineBasis` mirrors `/root/reference/uf3/representation/bspline.py:20-720`
attribute-for-attribute (knots_map, r_min_map, r_max_map, resolution_map,
symmetry, template_mask, flat_weights, templates, partition_sizes, col_idx,
frozen_c, r_cut, compress_3B/decompress_3B, column names, regularizer layout),
because downstream code (regression, model JSON files, LAMMPS export, column
dropping in `optimize.py`) addresses features through it.  It is pure set-up:
a few hundred doubles built once.  Nothing here evaluates splines over atoms —
that is the CUDA path (`uf3_b200/csrc/`), fed by `uf3_b200/tables.py`.

Defaults that pin results (reference file:line):
  pair  r_min 1.0, r_max 8.0, 15 intervals                 bspline.py:243-245
  trio  r_min [m,m,m], r_max [M,M,2M] (M default 4.0), [5,5,10]  :246-258
  trims leading {2:0, 3:3}, trailing {2:3, 3:3}             :66-67
  uniform knots rounded to 10 decimals, ends repeated x3    :1032-1035, :989
  r_cut = max(pair r_max, first two trio-leg r_max)         :188-202
"""
import itertools
import os
import re
import warnings
from typing import Collection, Dict, List, Tuple

import numpy as np

from uf3_b200 import composition, json_io, regularize

DEFAULT_LEADING_TRIM = {2: 0, 3: 3}
DEFAULT_TRAILING_TRIM = {2: 3, 3: 3}
_SCALAR = (float, np.floating, int, np.integer)


class BSplineBasis:
    def __init__(self, chemical_system, r_min_map=None, r_max_map=None,
                 resolution_map=None, knot_strategy="linear", offset_1b=True,
                 leading_trim=None, trailing_trim=None, knots_map=None):
        self.chemical_system = chemical_system
        self.knot_strategy = knot_strategy
        self.offset_1b = offset_1b
        self.leading_trim = process_trim_values(leading_trim, DEFAULT_LEADING_TRIM)
        self.trailing_trim = process_trim_values(trailing_trim, DEFAULT_TRAILING_TRIM)
        self.r_min_map = {}
        self.r_max_map = {}
        self.resolution_map = {}
        self.knots_map = {}
        self.knot_subintervals = {}
        self.symmetry = {}
        self.flat_weights = {}
        self.template_mask = {}
        self.templates = {}
        self.partition_sizes = []
        self.frozen_c = []
        self.col_idx = []
        self.r_cut = 0.0
        self._basis_functions = None
        self.update_knots(r_max_map, r_min_map, resolution_map, knots_map)
        self.knot_spacer = get_knot_spacer(self.knot_strategy)
        self.update_basis_functions()

    # ------------------------------------------------------------------ I/O
    @staticmethod
    def from_config(config):
        return BSplineBasis.from_dict(config)

    @staticmethod
    def from_dict(config):
        chemical_system = composition.ChemicalSystem.from_dict(config)
        settings = {}
        if "knots_path" in config and config.get("load_knots"):
            path = config["knots_path"]
            if os.path.isfile(path):
                try:
                    settings["knots_map"] = json_io.load_interaction_map(path)["knots"]
                except (ValueError, KeyError, IOError):
                    settings["knots_map"] = None
        for short, full in (("r_min", "r_min_map"), ("r_max", "r_max_map"),
                            ("resolution", "resolution_map"),
                            ("fit_offsets", "offset_1b")):
            if short in config:
                settings[full] = config[short]
            if full in config:
                settings[full] = config[full]
        for key in ("knot_strategy", "offset_1b", "leading_trim",
                    "trailing_trim", "knots_map"):
            if key in config:
                settings[key] = config[key]
        for key in ("leading_trim", "trailing_trim"):  # JSON stringifies int keys
            if isinstance(settings.get(key), dict):
                settings[key] = {int(k): v for k, v in settings[key].items()}
        basis = BSplineBasis(chemical_system, **settings)
        if "knots_path" in config and config.get("dump_knots"):
            json_io.dump_interaction_map(dict(knots=basis.knots_map),
                                         filename=config["knots_path"], write=True)
        return basis

    def as_dict(self):
        return dict(knot_strategy=self.knot_strategy,
                    offset_1b=self.offset_1b,
                    leading_trim={str(k): v for k, v in self.leading_trim.items()},
                    trailing_trim={str(k): v for k, v in self.trailing_trim.items()},
                    knots_map=self.knots_map,
                    **self.chemical_system.as_dict())

    # ----------------------------------------------------------- properties
    @property
    def degree(self):
        return self.chemical_system.degree

    @property
    def element_list(self):
        return self.chemical_system.element_list

    @property
    def interactions_map(self):
        return self.chemical_system.interactions_map

    @property
    def interactions(self):
        return self.chemical_system.interactions

    @property
    def n_feats(self) -> int:
        return int(np.sum(self.get_feature_partition_sizes()))

    @property
    def basis_functions(self):
        """scipy `BSpline.basis_element` callables per interaction (lazy).

        Kept for API compatibility (`bspline.py:77,337,362`); the CUDA path
        never touches these."""
        if self._basis_functions is None:
            built = {}
            for pair in self.interactions_map.get(2, []):
                built[pair] = generate_basis_functions(self.knot_subintervals[pair])
            if self.degree > 2:
                for trio in self.interactions_map.get(3, []):
                    built[trio] = [generate_basis_functions(sub)
                                   for sub in self.knot_subintervals[trio]]
            self._basis_functions = built
        return self._basis_functions

    def __repr__(self):
        sizes = self.get_interaction_partitions()[0]
        lines = ["BSplineBasis:", "    Basis functions:"]
        for n in range(2, self.degree + 1):
            for interaction in self.interactions_map[n]:
                lines.append(" " * 8 + f"{interaction}: {sizes[interaction]:d}")
        lines.append(repr(self.chemical_system))
        return "\n".join(lines)

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_basis_functions"] = None  # rebuilt lazily after unpickling
        return state

    # ---------------------------------------------------------------- knots
    def get_cutoff(self):
        reach = []
        for interaction, r_max in self.r_max_map.items():
            if isinstance(r_max, _SCALAR):
                reach.append(r_max)
            else:  # many-body: only legs that touch the centre atom
                reach.append(max(r_max[:len(interaction) - 1]))
        return max(reach)

    def update_knots(self, r_max_map=None, r_min_map=None, resolution_map=None,
                     knots_map=None):
        r_min_map = composition.sort_interaction_map(r_min_map or {})
        r_max_map = composition.sort_interaction_map(r_max_map or {})
        resolution_map = composition.sort_interaction_map(resolution_map or {})
        self.r_min_map.update(r_min_map)
        self.r_max_map.update(r_max_map)
        self.resolution_map.update(resolution_map)
        if knots_map is not None:
            self.update_knots_from_dict(composition.sort_interaction_map(knots_map))
        for map_ in (self.r_min_map, self.r_max_map, self.resolution_map):
            tuple_consistency_check(map_, self.interactions_map)
        for pair in self.interactions_map.get(2, []):
            self.r_min_map.setdefault(pair, 1.0)
            self.r_max_map.setdefault(pair, 8.0)
            self.resolution_map.setdefault(pair, 15)
        for trio in self.interactions_map.get(3, []):
            # Reference quirk kept (bspline.py:247-252): leg defaults are looked
            # up in the *argument* maps with unsorted combination keys.
            legs = list(itertools.combinations(trio, 2))
            low = np.min([r_min_map.get(k, 1.0) for k in legs])
            high = np.max([r_max_map.get(k, 4.0) for k in legs])
            self.r_min_map.setdefault(trio, [low, low, low])
            self.r_max_map.setdefault(trio, [high, high, 2 * high])
            self.resolution_map.setdefault(trio, [5, 5, 10])
            self.symmetry[trio] = find_symmetry_3B(trio,
                                                   self.r_min_map[trio],
                                                   self.r_max_map[trio],
                                                   self.resolution_map[trio])
        self.r_cut = self.get_cutoff()

    def update_knots_from_dict(self, knots_map):
        for pair in self.interactions_map.get(2, []):
            if pair not in knots_map:
                warnings.warn(f"{pair} specification unused.")
                continue
            seq = np.array(knots_map[pair])
            self.knots_map[pair] = seq
            self.r_min_map[pair] = seq[0]
            self.r_max_map[pair] = seq[-1]
            self.resolution_map[pair] = len(seq) - 7
        for trio in self.interactions_map.get(3, []):
            if trio not in knots_map:
                warnings.warn(f"{trio} specification unused.")
                continue
            given = knots_map[trio]
            if isinstance(given[0], _SCALAR):      # one sequence: all legs alike
                self.symmetry[trio] = 3
                legs = [given, given, given]
            elif len(given) == 2:                  # (l = m, n)
                self.symmetry[trio] = 2
                legs = [given[0], given[0], given[1]]
            else:
                if len(given) > 3:
                    warnings.warn("More than three knot sequences provided "
                                  f"for {trio} interaction.", RuntimeWarning)
                self.symmetry[trio] = 1
                legs = [given[0], given[1], given[2]]
            legs = [np.array(leg) for leg in legs]
            self.knots_map[trio] = legs
            self.r_min_map[trio] = [leg[0] for leg in legs]
            self.r_max_map[trio] = [leg[-1] for leg in legs]
            self.resolution_map[trio] = [len(leg) - 7 for leg in legs]

    def update_basis_functions(self):
        self._basis_functions = None
        for pair in self.interactions_map.get(2, []):
            if pair not in self.knots_map:
                seq = self.knot_spacer(self.r_min_map[pair], self.r_max_map[pair],
                                       self.resolution_map[pair])
                if self.r_min_map[pair] is None:
                    self.r_min_map[pair] = seq[0]
                self.knots_map[pair] = seq
            self.knot_subintervals[pair] = get_knot_subintervals(self.knots_map[pair])
        if self.degree > 2:
            for trio in self.interactions_map.get(3, []):
                if trio not in self.knots_map:
                    self.knots_map[trio] = [
                        self.knot_spacer(self.r_min_map[trio][leg],
                                         self.r_max_map[trio][leg],
                                         self.resolution_map[trio][leg])
                        for leg in range(3)]
                self.knot_subintervals[trio] = [get_knot_subintervals(seq)
                                                for seq in self.knots_map[trio]]
            self.set_flatten_template_3B()
        self.partition_sizes = self.get_feature_partition_sizes()
        self.col_idx, self.frozen_c = self.generate_frozen_indices(
            offset_1b=self.offset_1b, n_lead=self.leading_trim,
            n_trail=self.trailing_trim)

    # ------------------------------------------------------- feature layout
    def get_feature_partition_sizes(self) -> List:
        sizes = [1] * len(self.element_list)
        for degree in range(2, self.degree + 1):
            for interaction in self.interactions_map[degree]:
                if degree == 2:
                    sizes.append(self.resolution_map[interaction] + 3)
                elif degree == 3:
                    sizes.append(int(np.count_nonzero(
                        self.flat_weights[interaction] > 0)))
                else:
                    raise ValueError(
                        "Four-body terms and beyond are not yet implemented.")
        self.partition_sizes = sizes
        return sizes

    def get_interaction_partitions(self):
        sizes = self.get_feature_partition_sizes()
        starts = np.concatenate([[0], np.cumsum(sizes)])
        component_sizes = {}
        component_offsets = {}
        for j, interaction in enumerate(self.interactions):
            component_sizes[interaction] = sizes[j]
            component_offsets[interaction] = starts[j]
        return component_sizes, component_offsets

    def get_column_names(self):
        names = ["y"] + [f"n_{el}" for el in self.element_list]
        sizes = self.get_interaction_partitions()[0]
        for n in range(2, self.degree + 1):
            for interaction in self.interactions_map[n]:
                stem = "".join(interaction)
                names.extend(stem + str(i) for i in range(sizes[interaction]))
        return names

    def generate_frozen_indices(self, offset_1b=True, n_lead=None, n_trail=None,
                                value=0.0):
        """Columns pinned to `value` during the fit (trimmed basis functions).

        Follows the reference literally (`bspline.py:577-635`), including that
        3-body entries are indices *within* the trio's own partition."""
        n_lead = self.leading_trim if n_lead is None else n_lead
        n_trail = self.trailing_trim if n_trail is None else n_trail
        sizes, offsets = self.get_interaction_partitions()
        col_idx = []
        for pair in self.interactions_map.get(2, []):
            start, size = offsets[pair], sizes[pair]
            col_idx.extend(start + t for t in range(n_lead[2]))
            col_idx.extend(start + size - t for t in range(1, n_trail[2] + 1))
        for trio in self.interactions_map.get(3, []):
            edge = np.zeros_like(self.templates[trio])
            for t in range(n_lead[3]):
                edge[t, :, :] = 1
                edge[:, t, :] = 1
                edge[:, :, t] = 1
            for t in range(1, n_trail[3] + 1):
                edge[-t, :, :] = 1
                edge[:, -t, :] = 1
                edge[:, :, -t] = 1
            col_idx.extend(np.where(self.compress_3B(edge, trio) > 0)[0])
        frozen_c = [value] * len(col_idx)
        if not offset_1b:
            for j in range(len(self.element_list)):
                col_idx.insert(0, j)
                frozen_c.insert(0, 0)
        return np.array(col_idx, dtype=int), np.array(frozen_c)

    # -------------------------------------------------- 3-body compression
    def set_flatten_template_3B(self):
        for trio in self.interactions_map[3]:
            template = get_symmetry_weights(self.symmetry[trio],
                                            *self.knots_map[trio],
                                            self.leading_trim[3],
                                            self.trailing_trim[3])
            flat = template.flatten()
            keep, = np.where(flat > 0)
            self.template_mask[trio] = keep
            self.flat_weights[trio] = flat[keep]
            self.templates[trio] = template

    def _symmetrized(self, grid, interaction):
        order = self.symmetry[interaction]
        if order == 2:
            return grid + grid.transpose(1, 0, 2)
        if order == 3:
            return sum(grid.transpose(p) for p in itertools.permutations(range(3)))
        return grid

    def compress_3B(self, grid, interaction, fitting=True):
        grid = np.asarray(grid)
        if fitting:
            scale = self.flat_weights[interaction]
        else:
            scale = {1: 1.0, 2: 0.5, 3: 1 / 6}[self.symmetry[interaction]]
        folded = self._symmetrized(grid, interaction)
        return folded.flat[self.template_mask[interaction]] * scale

    def decompress_3B(self, vec, interaction):
        shape = tuple(len(seq) - 4 for seq in self.knots_map[interaction])
        grid = np.zeros(shape)
        grid.flat[self.template_mask[interaction]] = (
            np.asarray(vec) * self.flat_weights[interaction])
        return self._symmetrized(grid, interaction)

    # -------------------------------------------------------- regularizer
    def get_regularization_matrix(self, ridge_map=None, curvature_map=None,
                                  **kwargs):
        """Stacked ridge / curvature penalty (`bspline.py:371-429`).

        Strengths may also be passed as keywords such as `ridge_1b=1e-8`,
        `curvature_2b=1e-6` (first letter selects the kind, digits the order)."""
        ridge_map = dict(ridge_map or {})
        curvature_map = dict(curvature_map or {})
        for key, val in kwargs.items():
            order = int(re.sub("[^0-9]", "", key))
            if key.lower().startswith("r"):
                ridge_map[order] = float(val)
            elif key.lower().startswith("c"):
                curvature_map[order] = float(val)
        grid = regularize.DEFAULT_REGULARIZER_GRID
        ridge_map = {1: grid["ridge_1b"], 2: grid["ridge_2b"],
                     3: grid["ridge_3b"], **ridge_map}
        curvature_map = {1: 0.0, 2: grid["curve_2b"], 3: grid["curve_3b"],
                         **curvature_map}
        blocks = [self.get_regularization_matrix_1b(len(self.element_list),
                                                    ridge=ridge_map[1])]
        for degree in range(2, self.degree + 1):
            for interaction in self.interactions_map[degree]:
                if degree == 2:
                    build = self.get_regularization_matrix_2b
                elif degree == 3:
                    build = self.get_regularization_matrix_3b
                else:
                    raise ValueError(
                        "Four-body terms and beyond are not yet implemented.")
                blocks.append(build(interaction, ridge=ridge_map[degree],
                                    curvature=curvature_map[degree]))
        return regularize.combine_regularizer_matrices(blocks)

    def get_regularization_matrix_1b(self, n_elements, ridge):
        return regularize.get_ridge_penalty_matrix(n_elements) * np.sqrt(ridge)

    def get_regularization_matrix_2b(self, interaction, ridge, curvature):
        n = self.resolution_map[interaction] + 3
        matrix = regularize.get_ridge_penalty_matrix(n) * np.sqrt(ridge)
        if curvature > 0:
            bend = regularize.get_curvature_penalty_matrix_1D(n) * np.sqrt(curvature)
            matrix = np.vstack((matrix, bend))
        return matrix

    def get_regularization_matrix_3b(self, interaction, ridge, curvature):
        mask = self.template_mask[interaction]
        matrix = regularize.get_ridge_penalty_matrix(len(mask)) * np.sqrt(ridge)
        if curvature > 0:
            res = self.resolution_map[interaction]
            full = regularize.get_curvature_penalty_matrix_3D(
                res[0] + 3, res[1] + 3, res[2] + 3, flatten=False)
            bend = np.zeros((len(mask), len(mask)))
            for row, flat_idx in enumerate(mask):
                bend[row] = self.compress_3B(full[flat_idx], interaction)
            matrix = np.vstack((matrix, bend * np.sqrt(curvature)))
        return matrix


# ---------------------------------------------------------------- helpers
def find_symmetry_3B(trio: Tuple, r_min: List, r_max: List, resolution: List):
    """Permutational symmetry of a trio about its centre (`bspline.py:723-763`).

    1: no mirror plane; 2: j<->k interchangeable; 3: all three alike."""
    if trio[1] != trio[2]:
        return 1
    legs = list(zip(r_min, r_max, resolution))
    if legs[0] == legs[1] == legs[2]:
        return 3 if trio[0] == trio[1] else 2
    if legs[0] == legs[1]:
        return 2
    return 1


def get_symmetry_weights(symmetry, l_space, m_space, n_space, n_lead=0, n_trail=3):
    """Weight template over the full (L, M, N) grid (`angles.py:677-735`).

    Zero where the basis function is redundant under the trio's symmetry, can
    never satisfy the triangle inequality, or is trimmed; fractional on mirror
    planes so that symmetrised features are not double counted."""
    L, M, N = len(l_space) - 4, len(m_space) - 4, len(n_space) - 4
    li, mi, ni = np.meshgrid(np.arange(L), np.arange(M), np.arange(N),
                             indexing="ij")
    template = np.ones((L, M, N))
    if symmetry == 2:
        template[li == mi] = 0.5
        template[li > mi] = 0
    elif symmetry == 3:
        template[(li == ni) | (li == mi) | (mi == ni)] = 0.5
        template[(li > mi) | (mi > ni)] = 0
        template[(li == mi) & (li == ni)] = 1 / 6
    l_lo, l_hi = np.asarray(l_space)[li], np.asarray(l_space)[li + 4]
    m_lo, m_hi = np.asarray(m_space)[mi], np.asarray(m_space)[mi + 4]
    n_lo, n_hi = np.asarray(n_space)[ni], np.asarray(n_space)[ni + 4]
    impossible = ((l_hi + m_hi <= n_lo) | (l_hi + n_hi <= m_lo)
                  | (m_hi + n_hi <= l_lo))
    template[impossible] = 0
    for t in range(n_lead):
        template[t, :, :] = 0
        template[:, t, :] = 0
        template[:, :, t] = 0
    for t in range(1, n_trail + 1):
        template[-t, :, :] = 0
        template[:, -t, :] = 0
        template[:, :, -t] = 0
    return template


def get_knot_spacer(knot_strategy):
    try:
        return {"linear": generate_uniform_knots,
                "lammps": generate_lammps_knots,
                "geometric": generate_geometric_knots,
                "inverse": generate_inv_knots}[knot_strategy]
    except KeyError:
        raise ValueError("Invalid value of knot_strategy:", knot_strategy) from None


def knot_sequence_from_points(knot_points: Collection) -> np.ndarray:
    """Repeat both end points three more times (clamped cubic knot vector)."""
    pts = np.asarray(knot_points)
    return np.concatenate([np.repeat(pts[0], 3), pts, np.repeat(pts[-1], 3)])


def get_knot_subintervals(knots: np.ndarray) -> List:
    return [knots[i:i + 5] for i in range(len(knots) - 4)]


def generate_basis_functions(knot_subintervals):
    from scipy import interpolate
    return [interpolate.BSpline.basis_element(sub, extrapolate=False)
            for sub in knot_subintervals]


def _finish(points, sequence):
    return knot_sequence_from_points(points) if sequence else points


def generate_uniform_knots(r_min, r_max, n_intervals, sequence=True, offset=3):
    if r_min is None:
        r_min = -offset * (r_max - 0.0) / (n_intervals - offset)
    points = np.linspace(r_min, r_max, n_intervals + 1)
    return np.round(_finish(points, sequence), 10)


def _require_lower_bound(r_min):
    if r_min is None:
        raise ValueError(
            "Automatic lower-bound is WIP for this knot spacing scheme.")


def generate_inv_knots(r_min, r_max, n_intervals, sequence=True):
    _require_lower_bound(r_min)
    return _finish(np.linspace(1 / r_min, 1 / r_max, n_intervals + 1) ** -1, sequence)


def generate_geometric_knots(r_min, r_max, n_intervals, sequence=True):
    _require_lower_bound(r_min)
    return _finish(np.geomspace(r_min, r_max, n_intervals + 1), sequence)


def generate_lammps_knots(r_min, r_max, n_intervals, sequence=True):
    _require_lower_bound(r_min)
    points = np.linspace(r_min ** 2, r_max ** 2, n_intervals + 1) ** 0.5
    return _finish(points, sequence)


def parse_knots_file(filename: str, chemical_system) -> Dict:
    data = json_io.load_interaction_map(filename)
    knots_map = {}
    for d in range(2, chemical_system.degree + 1):
        for interaction in chemical_system.interactions_map[d]:
            if interaction not in data:
                continue
            seq = data[interaction]
            if (np.ptp(seq[:4]) == 0 and np.ptp(seq[-4:]) == 0
                    and np.all(np.gradient(seq) >= 0)):
                knots_map[interaction] = seq
    return knots_map


def tuple_consistency_check(map_, interaction_map):
    known = [i for group in interaction_map.values() for i in group]
    for entry in map_:
        if entry not in known:
            warnings.warn(f"{entry} specification unused.")


def process_trim_values(user_input, default_trim: Dict[int, int]):
    if user_input is None:
        return dict(default_trim)
    if isinstance(user_input, int):
        return {order: user_input for order in default_trim}
    if isinstance(user_input, dict):
        if not all(isinstance(k, int) for k in user_input):
            raise ValueError("Keys of the trimming values (order of interaction)"
                             " must be integers.")
        if not all(isinstance(v, int) for v in user_input.values()):
            raise ValueError("Values of the trimming values must be integers.")
        return dict(user_input)
    raise ValueError("Invalid input for trimming values. "
                     "Must be None, int, or a dict.")


def find_spline_indices(points, knot_sequence):
    """First of the four non-zero cubic basis functions at each point.

    A point exactly on a knot belongs to the interval on its left
    (`bspline.py:966`: `searchsorted(..., side='left') - 4`)."""
    return np.searchsorted(knot_sequence, points, side="left") - 4


def fit_spline_1d(x, y, knot_sequence):
    """Cubic B-spline coefficients of a sampled 1-D function in the least-squares sense — for comparing
    fitted pair potentials with a known curve, or for building a model from one
    (reference: representation/bspline.py:898-950, which hands the samples to FITPACK's
    LSQUnivariateSpline; here the same normal equations are solved from the B-spline design matrix).
    Samples outside the open knot range are dropped.  The reference then pads the samples so that no knot
    interval is empty, with a rule that is kept as it is because it shapes the result: for every knot
    interval whose LEFT edge lies below the smallest sample it adds (midpoint, y of the smallest sample) —
    which always includes the first interval — and for every interval whose left edge lies above the
    largest sample (midpoint, y of the largest sample)."""
    from scipy.interpolate import BSpline
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    t = np.asarray(knot_sequence, dtype=np.float64)
    inside = (x > t[0]) & (x < t[-1])
    x, y = x[inside], y[inside]
    if len(x) == 0:
        raise ValueError("no sample inside the knot range")
    lo, hi = int(np.argmin(x)), int(np.argmax(x))
    x_min, y_min, x_max, y_max = x[lo], y[lo], x[hi], y[hi]
    edges = np.unique(t)
    pad_x, pad_y = [], []
    for left, right in zip(edges[:-1], edges[1:]):
        if x_min > left:
            pad_x.append(0.5 * (left + right))
            pad_y.append(y_min)
        elif x_max < left:
            pad_x.append(0.5 * (left + right))
            pad_y.append(y_max)
    x, y = np.concatenate([x, pad_x]), np.concatenate([y, pad_y])
    order = np.argsort(x, kind="stable")
    design = BSpline.design_matrix(x[order], t, 3).toarray()
    coefficients, *_ = np.linalg.lstsq(design, y[order], rcond=None)
    return coefficients
