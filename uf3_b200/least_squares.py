"""`WeightedLinearModel` — model container and normal-equation fit.

API mirror of `/root/reference/uf3/regression/least_squares.py` for the parts the hot
path touches: loading / saving fitted models (`from_json`, `load`, `as_dict`, :185-216,
:528-621), the frozen-column bookkeeping (:817-890) and the weighted Gram solve
(`fit`, `fit_with_gram`, `combine_weighted_gram`, :248-353).  The Gram matrices can come
from host arrays (as in the reference) or straight from the device accumulator
(`GramAccumulator`, fed by `uf3b_gram_accumulate` on rows that never leave the GPU) —
the regularised solve itself is a p x p problem (p <= ~900) and stays on the host
LAPACK, exactly as the reference does it (`np.linalg.solve`, :763-771).
"""
import ctypes as C
import warnings
from typing import Dict

import numpy as np

from uf3_b200 import bspline, composition, json_io


# ---------------------------------------------------------------- helpers
def get_freezing_mask(n_feats, col_idx):
    return np.setdiff1d(np.arange(n_feats), col_idx)


def freeze_columns(x, y, mask, frozen_c, col_idx):
    """Drop the frozen columns of x and move their fixed contribution into y."""
    x = np.asarray(x)
    y = np.subtract(y, np.dot(x[:, col_idx], frozen_c))
    return x[:, mask], y


def freeze_regularizer(regularizer, mask):
    return regularizer[:, mask]


def revert_frozen_coefficients(solution, n_coeff, mask, frozen_c, frozen_idx):
    full = np.zeros(n_coeff, dtype=np.asarray(solution).dtype)
    full[np.asarray(mask, dtype=int)] = solution
    if len(frozen_idx):
        full[np.asarray(frozen_idx, dtype=int)] = frozen_c
    return full


def moore_penrose_components(x, y):
    x = np.asarray(x)
    return np.dot(x.T, x), np.dot(x.T, y)


def batched_moore_penrose(x, y, batch_size=2500):
    n_samples, n_features = np.shape(x)
    n_batches = int(n_samples / batch_size)
    if n_batches <= 1:
        return moore_penrose_components(x, y)
    gram = np.zeros((n_features, n_features))
    ordinate = np.zeros(n_features)
    for batch in np.array_split(np.arange(len(y)), n_batches):
        g, o = moore_penrose_components(x[batch], y[batch])
        gram += g
        ordinate += o
    return gram, ordinate


def dataframe_to_tuples(df_features, n_elements=None, energy_key="energy", sample_weights=None):
    """Feature DataFrame -> (x_e, y_e, x_f, y_f) (least_squares.py:666-716): first column is
    the target; energy rows (and targets) are divided by the atom count = sum of the leading
    `n_elements` composition columns; optional per-configuration sample weights."""
    names = df_features.index.get_level_values(0)
    energy_mask = df_features.index.get_level_values(-1) == energy_key
    force_mask = np.logical_not(energy_mask)
    data = df_features.to_numpy()
    y, x = data[:, 0], data[:, 1:]
    y_e, y_f = y[energy_mask], y[force_mask]
    if n_elements is not None:
        size = np.sum(x[energy_mask, :n_elements], axis=1)
        x_e = np.divide(x[energy_mask].T, size).T
        y_e = y_e / size
    else:
        x_e = x[energy_mask]
    x_f = x[force_mask]
    if sample_weights is not None:
        w = np.array([sample_weights.get(name, 1.0) for name in names])
        w_e, w_f = w[energy_mask], w[force_mask]
        x_e, y_e = np.multiply(x_e.T, w_e).T, np.multiply(y_e, w_e)
        x_f, y_f = np.multiply(x_f.T, w_f).T, np.multiply(y_f, w_f)
    return x_e, y_e, x_f, y_f


def lu_factorization(a, b):
    """Host LAPACK LU solve, as the reference (least_squares.py:763-771)."""
    return np.linalg.solve(a, b)


def device_solve(a, b, stream=None):
    """The same solve on the GPU: cuSOLVER getrf + getrs behind the C ABI (`uf3b_solve`)."""
    import ctypes as C
    from uf3_b200 import _native
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    if a.ndim != 2 or a.shape[0] != a.shape[1] or b.shape[0] != a.shape[0]:
        raise ValueError("device_solve expects a square matrix and a matching right-hand side")
    rhs = np.ascontiguousarray(b.T if b.ndim == 2 else b)          # [n_rhs][n]
    x = np.empty_like(rhs)
    _native.check(_native.lib().uf3b_solve(C.c_void_p(a.ctypes.data), C.c_void_p(rhs.ctypes.data),
                                           a.shape[0], 1 if b.ndim == 1 else b.shape[1],
                                           C.c_void_p(x.ctypes.data), stream))
    return x.T.copy() if b.ndim == 2 else x


class VarianceRecorder:
    """Running mean / standard deviation / count over batches (least_squares.py:19-68).  Batches are merged
    with the pairwise update of Chan et al., so the result does not depend on how the samples were cut."""

    def __init__(self, mean=0, std=0, n=0):
        self.mean, self.std, self.n = mean, std, int(n)

    def update(self, batch):
        batch = np.asarray(batch, dtype=np.float64)
        k = len(batch)
        if k == 0:
            return self.mean, self.std, self.n
        b_mean, b_var = np.mean(batch, axis=0), np.var(batch, axis=0)
        if self.n == 0:
            self.mean, self.std, self.n = b_mean, np.sqrt(b_var), k
            return self.mean, self.std, self.n
        m, total = float(self.n), float(self.n + k)
        var = m / total * self.std ** 2 + k / total * b_var + m * k / total ** 2 * (self.mean - b_mean) ** 2
        self.mean = (m * self.mean + k * b_mean) / total
        self.std = np.sqrt(var)
        self.n += k
        return self.mean, self.std, self.n

    def update_with_components(self, df, keys=None):
        """Force components of a data frame (columns fx, fy, fz holding scalars or per-atom arrays)."""
        values = []
        for row in df[list(keys or ("fx", "fy", "fz"))].itertuples(index=False):
            if any(component is np.nan for component in row):
                continue
            for component in row:
                values.extend(np.ravel(component))
        return self.update(values)


def apply_weights(x, y, weights=None):
    """Rows and targets scaled by per-sample weights (least_squares.py:892-913)."""
    if weights is None:
        return np.asarray(x), np.asarray(y)
    weights = np.asarray(weights, dtype=np.float64)
    if weights.shape != np.shape(y) or np.any(weights < 0):
        raise ValueError("weights must be non-negative, one per sample")
    return np.asarray(x) * weights[:, None], np.asarray(y) * weights


def validate_regularizer(regularizer, n_feats):
    n_row, n_col = np.shape(regularizer)
    if n_col != n_feats:
        raise ValueError(f"Expected regularizer shape: N x {n_feats}. Provided: {n_row} x {n_col}")


def linear_least_squares(x, y):
    """Normal-equation solve of x c = y (least_squares.py:774-787); a Tikhonov matrix is appended by the caller."""
    return lu_factorization(*moore_penrose_components(x, y))


def weighted_least_squares(x, y, weights=None, regularizer=None):
    """Weighted rows with an optional regularizer stacked below them (least_squares.py:790-814)."""
    x_fit, y_fit = apply_weights(x, y, weights)
    if regularizer is not None:
        x_fit = np.concatenate([x_fit, regularizer])
        y_fit = np.concatenate([y_fit, np.zeros(len(regularizer))])
    return linear_least_squares(x_fit, y_fit)


def subset_prediction(df, model, subset_keys=None, **kwargs):
    """(y_e, p_e, y_f, p_f) of the configurations `subset_keys` of a feature frame (least_squares.py:933-962)."""
    if subset_keys is not None:
        found = df.index.unique(level=0).intersection(subset_keys)
        if len(found) == 0:
            return [], [], [], []
        df = df.loc[found]
    x_e, y_e, x_f, y_f = dataframe_to_tuples(df, **kwargs)
    return y_e, model.predict(x_e), y_f, model.predict(x_f)


def batched_prediction(model, filename, table_names=None, subset_keys=None, drop_columns=None, **kwargs):
    """The same over the tables of a feature store (least_squares.py:965-1010)."""
    from uf3_b200 import store
    if table_names is None:
        _, _, table_names, _ = store.analyze_hdf_tables(filename)
    parts = [[], [], [], []]
    for df in store.dataframe_batch_loader(filename, table_names):
        if drop_columns is not None:
            df = df.drop(columns=drop_columns)
        for part, values in zip(parts, subset_prediction(df, model, subset_keys=subset_keys, **kwargs)):
            part.append(np.asarray(values))
    return tuple(np.concatenate(part) if part else np.zeros(0) for part in parts)


def calc_E_F_weights(n_e, n_f, std_e, std_f):
    """Weights of the energy / force blocks (least_squares.py:1147-1168)."""
    if std_e == 0:
        return 1.0, 1 / np.sqrt(n_f)
    return 1 / np.sqrt(n_e) / std_e, 1 / np.sqrt(n_f) / std_f


def arrange_coefficients(coefficients, bspline_config):
    """Flat coefficient vector -> {element: float, pair: vector, trio: vector}."""
    pieces = np.array_split(coefficients, np.cumsum(bspline_config.partition_sizes)[:-1])
    elements = bspline_config.element_list
    solutions = {el: piece[0] for el, piece in zip(elements, pieces[:len(elements)])}
    rest = pieces[len(elements):]
    j = 0
    for degree in range(2, bspline_config.degree + 1):
        for interaction in bspline_config.interactions_map[degree]:
            solutions[interaction] = rest[j]
            j += 1
    return solutions


def rmse_metric(predicted, actual):
    return float(np.sqrt(np.mean(np.subtract(predicted, actual) ** 2)))


def mae_metric(predicted, actual):
    return float(np.mean(np.abs(np.subtract(predicted, actual))))


# ---------------------------------------------------------------- model
class WeightedLinearModel:
    def __init__(self, bspline_config, regularizer=None, data_coverage=None, solver="host", **params):
        self.coefficients = None
        self.solver = solver        # "host" = numpy LAPACK as the reference; "cusolver" = on the GPU
        self.regularizer = regularizer
        self.bspline_config = bspline_config
        n_basis = int(np.sum(bspline_config.get_feature_partition_sizes()))
        if data_coverage is not None:
            if len(data_coverage) != n_basis:
                raise ValueError(f"Incorrect data_coverage shape: {len(data_coverage)} != {n_basis}")
            self.data_coverage = np.asarray(data_coverage)
        else:
            self.data_coverage = np.zeros(n_basis, dtype=bool)
        if self.regularizer is None:
            self.set_params(**params)

    def set_params(self, **params):
        if "bspline_config" in params:
            self.bspline_config = params["bspline_config"]
        if "regularizer" in params:
            self.regularizer = params["regularizer"]
        elif self.regularizer is None:
            strengths = {k: v for k, v in params.items()
                         if isinstance(v, (int, float, np.floating))}
            self.regularizer = self.bspline_config.get_regularization_matrix(**strengths)

    # ------------------------------------------------------------ I/O
    @staticmethod
    def from_config(config):
        return WeightedLinearModel.from_dict(config)

    @staticmethod
    def from_dict(config):
        basis = bspline.BSplineBasis.from_dict(config)
        model = WeightedLinearModel(basis, regularizer=config.get("regularizer"),
                                    data_coverage=config.get("data_coverage"))
        model.load(solution=config)
        return model

    @staticmethod
    def from_json(filename):
        return WeightedLinearModel.from_dict(json_io.load_interaction_map(filename))

    def as_dict(self):
        solution = arrange_coefficients(self.coefficients, self.bspline_config)
        for trio in self.bspline_config.interactions_map.get(3, []):
            solution[trio] = self.bspline_config.decompress_3B(solution[trio], trio)
        return dict(coefficients=solution, knots=self.bspline_config.knots_map,
                    data_coverage=self.data_coverage, **self.bspline_config.as_dict())

    dump = as_dict

    def to_json(self, filename):
        json_io.dump_interaction_map(self.as_dict(), filename=filename, write=True)

    def load(self, solution: Dict = None, filename: str = None):
        """Flatten {interaction: coefficients} into `self.coefficients`; 3-body entries may
        be full (L, M, N) grids, which are folded with `compress_3B(fitting=False)`."""
        if filename is not None:
            if solution is not None:
                warnings.warn("Provided solutions ignored; loading file.")
            solution = json_io.load_interaction_map(filename)
        elif solution is None:
            raise ValueError("Neither solution nor filename were provided.")
        if "coefficients" in solution:
            solution = solution["coefficients"]
        elif "solution" in solution:
            warnings.warn("'solution' should be renamed to 'coefficients'")
            solution = solution["solution"]
        solution = dict(solution)
        for key in list(solution):
            if isinstance(key, tuple):
                solution.setdefault(composition.sort_interaction_symbols(key), solution[key])
        basis = self.bspline_config
        sizes = basis.get_interaction_partitions()[0]
        flat = [[solution[el]] for el in basis.element_list]
        for pair in basis.interactions_map[2]:
            if pair not in solution:
                warnings.warn(f"{pair} not provided.")
                solution[pair] = np.zeros(sizes[pair])
            if len(solution[pair]) != sizes[pair]:
                raise ValueError(f"Incorrect shape: {pair}, {len(solution[pair])} != {sizes[pair]}")
            flat.append(np.asarray(solution[pair], dtype=float))
        for trio in (basis.interactions_map.get(3, []) if basis.degree > 2 else []):
            if trio not in solution:
                raise ValueError(f"{trio} not provided.")
            component = np.array(solution[trio])
            if component.ndim > 1:
                component = basis.compress_3B(component, trio, fitting=False)
            if len(component) != sizes[trio]:
                raise ValueError(f"Incorrect shape: {trio}, {len(component)} != {sizes[trio]}")
            flat.append(np.asarray(component, dtype=float))
        flat = np.concatenate(flat)
        if len(flat) != int(np.sum(basis.partition_sizes)):
            raise ValueError(f"Incorrect coefficients: {len(flat)} provided, "
                             f"{int(np.sum(basis.partition_sizes))} expected.")
        self.coefficients = flat

    # ------------------------------------------------------------ bookkeeping
    @property
    def n_feats(self):
        return self.bspline_config.n_feats

    @property
    def frozen_c(self):
        return self.bspline_config.frozen_c

    @property
    def col_idx(self):
        return self.bspline_config.col_idx

    @property
    def mask(self):
        # cached per (n_feats, frozen columns): the fit at the end of a stream of frames reads it several
        # times and np.setdiff1d costs 0.15 ms a call
        key = (int(self.n_feats), tuple(int(c) for c in np.atleast_1d(self.col_idx)))
        cached = getattr(self, "_mask_cache", None)
        if cached is None or cached[0] != key:
            cached = (key, get_freezing_mask(self.n_feats, self.col_idx))
            self._mask_cache = cached
        return cached[1]

    def _regularizer_gram(self, mask):
        """R^T R on the unfrozen columns, cached while the regularizer object and the mask stay the same."""
        cached = getattr(self, "_reg_cache", None)
        if cached is None or cached[0] is not self.regularizer or cached[1] is not mask:
            reg = freeze_regularizer(self.regularizer, mask)
            cached = (self.regularizer, mask, np.dot(reg.T, reg))
            self._reg_cache = cached
        return cached[2]

    def __repr__(self):
        return "\n".join(["WeightedLinearModel:", f"    Fit: {self.coefficients is not None}",
                          repr(self.bspline_config)])

    # ------------------------------------------------------------ fit
    def fit_with_gram(self, gram, ordinate):
        """Solve (G + R^T R) c = b on the unfrozen columns (least_squares.py:248-272)."""
        coverage = revert_frozen_coefficients(np.sum(gram, axis=0) != 0, self.n_feats, self.mask,
                                              self.frozen_c, self.col_idx)
        self.data_coverage = np.logical_or(self.data_coverage, coverage)
        solve = device_solve if getattr(self, "solver", "host") == "cusolver" else lu_factorization
        solution = solve(gram + self._regularizer_gram(self.mask), ordinate)
        self.coefficients = revert_frozen_coefficients(solution, self.n_feats, self.mask,
                                                       self.frozen_c, self.col_idx)

    def combine_weighted_gram(self, gram_e, gram_f, ord_e, ord_f, energy_weight, force_weight, weight):
        gram = weight * energy_weight ** 2 * gram_e + (1 - weight) * force_weight ** 2 * gram_f
        ordinate = weight * energy_weight ** 2 * ord_e + (1 - weight) * force_weight ** 2 * ord_f
        return gram, ordinate

    def fit(self, x_e, y_e, x_f=None, y_f=None, weight=0.5, batch_size=2500):
        x_e, y_e = freeze_columns(x_e, y_e, self.mask, self.frozen_c, self.col_idx)
        gram, ordinate = batched_moore_penrose(x_e, y_e, batch_size=batch_size)
        if x_f is not None:
            w_e, w_f = calc_E_F_weights(len(y_e), len(y_f), np.std(y_e), np.std(y_f))
            x_f, y_f = freeze_columns(x_f, y_f, self.mask, self.frozen_c, self.col_idx)
            gram_f, ord_f = batched_moore_penrose(x_f, y_f, batch_size=batch_size)
            gram, ordinate = self.combine_weighted_gram(gram, gram_f, ordinate, ord_f, w_e, w_f, weight)
        self.fit_with_gram(gram, ordinate)

    def fit_from_accumulator(self, acc, weight=0.5):
        """Fit from full-width Gram statistics gathered on the device (`GramAccumulator`,
        all-reduced over ranks by `uf3_b200.distributed`).  Equivalent to `fit` on the
        stacked rows: frozen columns are eliminated from the Gram blocks instead of from X."""
        stats = acc.export()
        mask, col, c0 = self.mask, np.asarray(self.col_idx, dtype=int), np.asarray(self.frozen_c)

        def reduce(gram, ordinate):
            # (X_m)^T X_m and (X_m)^T (y - X_c c0)
            return gram[np.ix_(mask, mask)], ordinate[mask] - gram[np.ix_(mask, col)] @ c0

        gram, ordinate = reduce(stats["gram_e"], stats["ord_e"])
        if stats["n_f"] > 0:
            gram_f, ord_f = reduce(stats["gram_f"], stats["ord_f"])
            # the reference takes np.std AFTER folding the frozen contribution into y;
            # frozen coefficients are zero (bspline.py:577-635), so y is unchanged
            w_e, w_f = calc_E_F_weights(stats["n_e"], stats["n_f"], stats["std_e"], stats["std_f"])
            gram, ordinate = self.combine_weighted_gram(gram, gram_f, ordinate, ord_f, w_e, w_f, weight)
        self.fit_with_gram(gram, ordinate)

    def initialize_gram_ordinate(self):
        """Zero Gram blocks and ordinates over the unfrozen columns (least_squares.py:425-433)."""
        p = self.n_feats - len(self.col_idx)
        return np.zeros((p, p)), np.zeros((p, p)), np.zeros(p), np.zeros(p)

    def gram_from_df(self, df, keys, e_variance=None, f_variance=None, sample_weights=None, energy_key="energy",
                     batch_size=2500):
        """(gram_e, gram_f, ord_e, ord_f) of the configurations `keys` of a feature frame, frozen columns
        eliminated; the variance recorders, if given, see the targets (least_squares.py:435-483)."""
        x_e, y_e, x_f, y_f = dataframe_to_tuples(df.loc[keys], n_elements=len(self.bspline_config.element_list),
                                                 energy_key=energy_key, sample_weights=sample_weights)
        x_e, y_e = freeze_columns(x_e, y_e, self.mask, self.frozen_c, self.col_idx)
        x_f, y_f = freeze_columns(x_f, y_f, self.mask, self.frozen_c, self.col_idx)
        if e_variance is not None and f_variance is not None:
            e_variance.update(y_e)
            f_variance.update(y_f)
        gram_e, ord_e = batched_moore_penrose(x_e, y_e, batch_size=batch_size)
        gram_f, ord_f = batched_moore_penrose(x_f, y_f, batch_size=batch_size)
        return gram_e, gram_f, ord_e, ord_f

    def fit_from_file(self, filename, subset, weight=0.5, batch_size=2500, sample_weights=None,
                      energy_key="energy", progress="bar", drop_columns=None, gram="auto"):
        """Fit from a chunked feature store written by `BasisFeaturizer.batched_to_hdf`
        (least_squares.py:355-433): the tables are read in sorted order, the rows of the
        configurations in `subset` are folded into the energy / force normal equations together
        with the running target statistics, and the regularised system is solved once.
        `gram`: "device" folds the force rows on the GPU (`uf3b_gram_accumulate` reads the host
        rows of a chunk; FP64 tensor cores), "host" uses numpy as the reference does, "auto" takes
        the device when the model was created with solver="cusolver"."""
        import os
        from uf3_b200 import store
        if not os.path.isfile(filename):
            raise FileNotFoundError(filename)
        on_device = gram == "device" or (gram == "auto" and getattr(self, "solver", "host") == "cusolver")
        n_tables, _, table_names, _ = store.analyze_hdf_tables(filename)
        n_elements = len(self.bspline_config.element_list)
        stats = None
        subset = list(subset)
        for table_name in table_names:
            df = store.load_feature_db(filename, table_name)
            keys = df.index.unique(level=0).intersection(subset)
            if len(keys) == 0:
                continue
            if drop_columns is not None:
                df = df.drop(columns=drop_columns)
            x_e, y_e, x_f, y_f = dataframe_to_tuples(df.loc[keys], n_elements=n_elements, energy_key=energy_key,
                                                     sample_weights=sample_weights)
            if stats is None:
                stats = (GramAccumulator if on_device else GramStats)(x_e.shape[1])
            stats.add_rows(x_e, y_e, is_force=False)
            if len(y_f):
                stats.add_rows(x_f, y_f, is_force=True)
        if stats is None:
            raise ValueError("no configuration of `subset` is in " + str(filename))
        self.fit_from_accumulator(stats, weight=weight)
        if on_device:
            stats.close()

    def batched_predict(self, filename, keys=None, table_names=None, score=True, drop_columns=None):
        """Targets and predictions of the rows of a feature store (least_squares.py:486-526, :1060-1118)."""
        from uf3_b200 import store
        if table_names is None:
            _, _, table_names, _ = store.analyze_hdf_tables(filename)
        n_elements = len(self.bspline_config.element_list)
        y_e, p_e, y_f, p_f = [], [], [], []
        for df in store.dataframe_batch_loader(filename, table_names):
            if keys is not None:
                found = df.index.unique(level=0).intersection(keys)
                if len(found) == 0:
                    continue
                df = df.loc[found]
            if drop_columns is not None:
                df = df.drop(columns=drop_columns)
            x_e, t_e, x_f, t_f = dataframe_to_tuples(df, n_elements=n_elements)
            y_e.append(t_e); p_e.append(self.predict(x_e))
            y_f.append(t_f); p_f.append(self.predict(x_f))
        y_e, p_e, y_f, p_f = (np.concatenate(v) if v else np.zeros(0) for v in (y_e, p_e, y_f, p_f))
        if score:
            return y_e, p_e, y_f, p_f, rmse_metric(y_e, p_e), rmse_metric(y_f, p_f)
        return y_e, p_e, y_f, p_f

    def predict(self, x):
        return np.dot(x, self.coefficients)

    def score(self, x, y):
        return rmse_metric(self.predict(x), y)


# ---------------------------------------------------------------- Gram statistics
class GramStats:
    """Normal-equation statistics of a stream of feature rows (host side).

    Holds G_e, b_e (energy rows), G_f, b_f (force rows) at FULL feature width and the
    running count / sum / sum of squares of the targets that replace the reference's
    `VarianceRecorder` (least_squares.py:19-53).  Energy rows and their targets are
    divided by the atom count on entry, as `dataframe_to_tuples` does (:697-700).
    `to_vector` / `from_vector` flatten everything into one float64 array of
    2 F^2 + 2 F + 6 values, so a fit sharded over many ranks costs ONE all-reduce."""

    def __init__(self, n_feats):
        n = self.n_feats = int(n_feats)
        self.gram_e, self.gram_f = np.zeros((n, n)), np.zeros((n, n))
        self.ord_e, self.ord_f = np.zeros(n), np.zeros(n)
        self.moments = np.zeros(6)      # n_e, sum_e, sumsq_e, n_f, sum_f, sumsq_f

    def _count(self, y, is_force):
        k = 3 if is_force else 0
        self.moments[k:k + 3] += (len(y), float(np.sum(y)), float(np.dot(y, y)))

    def add_energy_row(self, x_energy, energy, n_atoms):
        x = np.asarray(x_energy, dtype=np.float64) / n_atoms
        y = float(energy) / n_atoms
        self.gram_e += np.outer(x, x)
        self.ord_e += x * y
        self._count(np.array([y]), False)

    def add_force_rows(self, x_forces, y_forces):
        x = np.asarray(x_forces, dtype=np.float64)
        y = np.asarray(y_forces, dtype=np.float64).reshape(-1)
        self.gram_f += x.T @ x
        self.ord_f += x.T @ y
        self._count(y, True)

    def add_rows(self, x, y, is_force):
        """Rows that are already in fit form (energy rows divided by the atom count)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
        if is_force:
            return self.add_force_rows(x, y)
        self.gram_e += x.T @ x
        self.ord_e += x.T @ y
        self._count(y, False)

    def _blocks(self):
        return self.gram_e, self.gram_f, self.ord_e, self.ord_f

    def to_vector(self):
        gram_e, gram_f, ord_e, ord_f = self._blocks()
        return np.concatenate([gram_e.ravel(), gram_f.ravel(), ord_e, ord_f, self.moments])

    def from_vector(self, vec):
        """Replace the state by a flattened one (e.g. the all-reduced sum over ranks)."""
        n, g = self.n_feats, self.n_feats ** 2
        vec = np.asarray(vec, dtype=np.float64)
        if vec.shape != (2 * g + 2 * n + 6,):
            raise ValueError("flattened Gram statistics have the wrong length")
        self._reset_device()
        self.gram_e, self.gram_f = vec[:g].reshape(n, n).copy(), vec[g:2 * g].reshape(n, n).copy()
        self.ord_e, self.ord_f = vec[2 * g:2 * g + n].copy(), vec[2 * g + n:2 * g + 2 * n].copy()
        self.moments = vec[2 * g + 2 * n:].copy()

    def _reset_device(self):
        pass

    def export(self):
        gram_e, gram_f, ord_e, ord_f = self._blocks()
        n_e, s_e, ss_e, n_f, s_f, ss_f = self.moments

        def std(count, total, squares):
            if count < 1:
                return 0.0
            return float(np.sqrt(max(squares / count - (total / count) ** 2, 0.0)))

        return dict(gram_e=gram_e, gram_f=gram_f, ord_e=ord_e, ord_f=ord_f,
                    n_e=int(round(n_e)), n_f=int(round(n_f)),
                    std_e=std(n_e, s_e, ss_e), std_f=std(n_f, s_f, ss_f))


class GramAccumulator(GramStats):
    """`GramStats` whose force block is accumulated ON THE DEVICE (C ABI `uf3b_gram_*`)
    from rows that the featurize kernel left in HBM: a frame's 3N x F force rows (110 MB
    for 10k atoms with the 456-column basis) are never copied to the host.  The energy
    row is one vector per frame and is added on the host."""

    def __init__(self, n_feats):
        super().__init__(n_feats)
        from uf3_b200 import _native
        self._native = _native
        self._lib = _native.lib()
        self._handle = C.c_void_p()
        _native.check(self._lib.uf3b_gram_create(self.n_feats, C.byref(self._handle)))

    def close(self):
        if getattr(self, "_handle", None):
            self._lib.uf3b_gram_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_force_rows_device(self, x_ptr, y_forces, rows, ld, stream=None, y_moments=None):
        """x_ptr: device address of `rows` x n_feats float64 rows (row stride ld doubles);
        y_forces: host targets in the same row order (fx_0.., fy_0.., fz_0..) — or, with
        `y_moments` = (sum, sum of squares) supplied by the caller, the DEVICE address of the
        targets, so that nothing crosses PCIe per frame."""
        if y_moments is not None:
            self._native.check(self._lib.uf3b_gram_accumulate(
                self._handle, C.c_void_p(x_ptr), C.c_void_p(int(y_forces)), int(rows), int(ld), 1, stream))
            self.moments[3:6] += (rows, float(y_moments[0]), float(y_moments[1]))
            return
        y = np.ascontiguousarray(y_forces, dtype=np.float64).reshape(-1)
        if len(y) != rows:
            raise ValueError("one target per force row is required")
        self._native.check(self._lib.uf3b_gram_accumulate(
            self._handle, C.c_void_p(x_ptr), C.c_void_p(y.ctypes.data), int(rows), int(ld), 1, stream))
        self._count(y, True)

    def add_force_rows(self, x_forces, y_forces):
        """Host rows folded by the device kernel (the C ABI takes host or device pointers)."""
        x = np.ascontiguousarray(x_forces, dtype=np.float64)
        y = np.ascontiguousarray(y_forces, dtype=np.float64).reshape(-1)
        if x.ndim != 2 or x.shape[1] != self.n_feats or len(y) != len(x):
            raise ValueError("force rows must be (rows, n_feats) with one target per row")
        self._native.check(self._lib.uf3b_gram_accumulate(
            self._handle, C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), len(y), x.shape[1], 1, None))
        self._count(y, True)

    def _device_force_block(self):
        n = self.n_feats
        gram, ordinate = np.zeros((n, n)), np.zeros(n)
        self._native.check(self._lib.uf3b_gram_export(
            self._handle, 1, C.c_void_p(gram.ctypes.data), C.c_void_p(ordinate.ctypes.data)))
        return gram, ordinate

    def _blocks(self):
        gram_d, ord_d = self._device_force_block()
        return self.gram_e, self.gram_f + gram_d, self.ord_e, self.ord_f + ord_d

    def _reset_device(self):
        self.close()
        self._native.check(self._lib.uf3b_gram_create(self.n_feats, C.byref(self._handle)))
