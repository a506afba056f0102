"""`BasisFeaturizer` — the fit-path entry point, API-compatible with the reference
(`/root/reference/uf3/representation/process.py:20-506`), computed by the CUDA path.

What is kept from the reference interface (same names, argument meaning, return layout
and error behaviour):
  evaluate_configuration(geom, name, energy, forces, energy_key) -> dict      process.py:293
      energy row  = [E, n_el..., 2-body, 3-body]; force rows fx_0..fx_{N-1}, fy_0.., fz_0..
      keys (name, 'energy') / (name, 'fx_3'); input forces are (3, N);
      unknown element -> RuntimeWarning and {}                               (:322-330)
  featurize_energy_2B/3B(geom, supercell=None) -> (F2,) / (F3,)              process.py:369,440
  featurize_force_2B/3B(geom, supercell=None)  -> (N, 3, F2) / (N, 3, F3)    process.py:402,473
      as in the reference, `supercell=None` means "no periodic images"
  evaluate(df, ...), evaluate_parallel(df, client, ...), arrange_features_dataframe
  columns, and the basis pass-through properties

What changes underneath: no dense distance matrices; one neighbour-list kernel and one
fused feature kernel per configuration (`uf3_b200/csrc`).  There is no CPU fallback: the
methods raise if libuf3b.so or a CUDA device is missing.
"""
import warnings

import numpy as np

from uf3_b200 import geometry
from uf3_b200.atoms import frame_arrays


class BasisFeaturizer:
    def __init__(self, bspline_config, fit_forces=True, prefix="x", device=None):
        self.bspline_config = bspline_config
        self.fit_forces = fit_forces
        self.prefix = prefix
        self.device = device
        self.columns = self.bspline_config.get_column_names()
        self._engine = None

    # ------------------------------------------------------------------ plumbing
    @staticmethod
    def from_config(bspline_config, config):
        keys = ("prefix", "fit_forces", "device")
        return BasisFeaturizer(bspline_config, **{k: v for k, v in config.items() if k in keys})

    def __getstate__(self):
        # the reference pickles featurizers into worker processes (util/parallel.py:182):
        # device handles stay behind and are re-created where the copy is used
        state = dict(self.__dict__)
        state["_engine"] = None
        return state

    @property
    def engine(self):
        if self._engine is None:
            from uf3_b200.engine import Engine
            self._engine = Engine(self.bspline_config, device=self.device)
        return self._engine

    def __repr__(self):
        return "\n".join(["BasisFeaturizer:", f"    Fit forces: {self.fit_forces}",
                          f"    Column prefix: {self.prefix}", repr(self.bspline_config)])

    chemical_system = property(lambda self: self.bspline_config.chemical_system)
    degree = property(lambda self: self.bspline_config.degree)
    element_list = property(lambda self: self.bspline_config.element_list)
    interactions_map = property(lambda self: self.bspline_config.interactions_map)
    r_min_map = property(lambda self: self.bspline_config.r_min_map)
    r_max_map = property(lambda self: self.bspline_config.r_max_map)
    resolution_map = property(lambda self: self.bspline_config.resolution_map)
    r_cut = property(lambda self: self.bspline_config.r_cut)
    knots_map = property(lambda self: self.bspline_config.knots_map)
    knot_subintervals = property(lambda self: self.bspline_config.knot_subintervals)
    basis_functions = property(lambda self: self.bspline_config.basis_functions)
    partition_sizes = property(lambda self: self.bspline_config.partition_sizes)
    interaction_hashes = property(lambda self: self.chemical_system.interaction_hashes)
    leading_trim = property(lambda self: self.bspline_config.leading_trim)
    trailing_trim = property(lambda self: self.bspline_config.trailing_trim)

    # ------------------------------------------------------------------ one configuration
    def _images(self, geom, supercell):
        """Periodic-image table for a call that passes `supercell` the reference's way."""
        positions, numbers, cell, pbc = frame_arrays(geom)
        if supercell is None or supercell is geom:
            return positions, numbers, None
        n = len(positions)
        sup = np.asarray(supercell.get_positions(), dtype=np.float64)
        if len(sup) == n:
            return positions, numbers, None
        if len(sup) % n or not np.array_equal(np.asarray(supercell.get_atomic_numbers())[:n], numbers):
            raise ValueError("supercell must be geometry.get_supercell(geom, r_cut) of the same geom")
        images = geometry.image_table(cell, pbc, self.r_cut)
        expect = (positions[None, :, :] + images[1][:, None, :]).reshape(-1, 3)
        if expect.shape != sup.shape or not np.allclose(expect, sup, rtol=0, atol=1e-9):
            raise ValueError("supercell does not match the periodic images of geom at r_cut")
        return positions, numbers, images

    def _rows(self, geom, images="auto", energy=True, forces=True):
        positions, numbers, cell, pbc = frame_arrays(geom)
        if isinstance(images, str):
            images = geometry.image_table(cell, pbc, self.r_cut) if np.any(pbc) else None
        eng = self.engine
        eng.build_neighbors(positions, numbers, images=images)
        return eng.featurize(energy=energy, forces=forces)

    def evaluate_configuration(self, geom, name=None, energy=None, forces=None, energy_key="energy"):
        eval_map = {}
        n_atoms = len(geom)
        invalid = set(geom.get_chemical_symbols()).difference(self.element_list)
        if invalid:
            message = "Invalid elements: {}".format(", ".join(sorted(invalid)))
            if name is not None:
                message += " in configuration " + str(name)
            warnings.warn(message, RuntimeWarning)
            return dict()
        want_e, want_f = energy is not None, forces is not None
        if not (want_e or want_f):
            return eval_map
        xe, xf = self._rows(geom, energy=want_e, forces=want_f)
        if want_e:
            key = (name, energy_key) if name is not None else energy_key
            eval_map[key] = np.insert(xe, 0, energy)
        if want_f:
            forces = np.asarray(forces, dtype=np.float64)
            if forces.shape != (3, n_atoms):
                raise ValueError("forces must have shape (3, n_atoms)")
            rows = np.empty((3 * n_atoms, xf.shape[1] + 1))
            rows[:, 0] = forces.reshape(-1)
            rows[:, 1:] = xf
            for j, component in enumerate(("fx", "fy", "fz")):
                for i in range(n_atoms):
                    atom_index = f"{component}_{i}"
                    key = (name, atom_index) if name is not None else atom_index
                    eval_map[key] = rows[j * n_atoms + i]
        return eval_map

    def _partition(self, degree):
        sizes, offsets = self.bspline_config.get_interaction_partitions()
        keys = self.interactions_map[degree]
        start = offsets[keys[0]]
        stop = offsets[keys[-1]] + sizes[keys[-1]]
        return int(start), int(stop)

    def _energy_part(self, geom, supercell, degree):
        positions, numbers, images = self._images(geom, supercell)
        self.engine.build_neighbors(positions, numbers, images=images)
        xe, _ = self.engine.featurize(energy=True, forces=False)
        a, b = self._partition(degree)
        return xe[a:b].copy()

    def _force_part(self, geom, supercell, degree):
        positions, numbers, images = self._images(geom, supercell)
        self.engine.build_neighbors(positions, numbers, images=images)
        _, xf = self.engine.featurize(energy=False, forces=True)
        a, b = self._partition(degree)
        n = len(positions)
        return np.ascontiguousarray(xf[:, a:b].reshape(3, n, b - a).transpose(1, 0, 2))

    def featurize_energy_2B(self, geom, supercell=None):
        return self._energy_part(geom, supercell, 2)

    def featurize_force_2B(self, geom, supercell=None):
        return self._force_part(geom, supercell, 2)

    def featurize_energy_3B(self, geom, supercell=None):
        return self._energy_part(geom, supercell, 3)

    def featurize_force_3B(self, geom, supercell=None):
        return self._force_part(geom, supercell, 3)

    # ------------------------------------------------------------------ data frames
    def evaluate(self, df_data, atoms_key="geometry", energy_key="energy", progress="bar"):
        """DataFrame of configurations -> DataFrame of feature rows (process.py:121-174)."""
        eval_map = {}
        header = list(df_data.columns)
        position = {key: header.index(key) + 1
                    for key in (atoms_key, energy_key, "fx", "fy", "fz") if key in header}
        for row in df_data.itertuples(name=None):
            name = row[0]
            geom = row[position[atoms_key]]
            energy = row[position[energy_key]] if energy_key in position else None
            forces = None
            if "fx" in position and self.fit_forces:
                forces = [row[position[c]] for c in ("fx", "fy", "fz")]
                if np.any(np.isnan(np.asarray(forces, dtype=float))):
                    forces = None
            eval_map.update(self.evaluate_configuration(geom, name, energy, forces, energy_key))
        return self.arrange_features_dataframe(eval_map)

    def arrange_features_dataframe(self, eval_map):
        import pandas as pd
        df = pd.DataFrame.from_dict(eval_map, orient="index", columns=self.columns)
        return df.set_index(pd.MultiIndex.from_tuples(df.index))

    def _visible_devices(self):
        import ctypes as C
        from uf3_b200 import _native
        count = C.c_int32()
        _native.check(_native.lib().uf3b_device_count(C.byref(count)))
        return int(count.value)

    def _on_device(self, device):
        """A copy of this featurizer bound to `device` (its engine is created where it is used)."""
        clone = BasisFeaturizer(self.bspline_config, fit_forces=self.fit_forces, prefix=self.prefix,
                                device=device)
        return clone

    def evaluate_parallel(self, df_data, client=None, atoms_key="geometry", energy_key="energy",
                          n_jobs=2, shuffle=True, progress="bar"):
        """Reference signature and semantics (process.py:196-254): the data frame is cut into
        `n_jobs` batches (shuffled first when `shuffle`), the batches are evaluated concurrently and
        the result is returned in the order of `df_data`.  The reference deals the batches over the
        worker processes of `client`; here batch k runs on visible GPU k mod n_gpus, each with its own
        engine.  `client` may be a `concurrent.futures` executor (a process pool receives pickled
        featurizers, which re-create their device handles in the worker, as the reference's do);
        with `client=None` one thread per GPU drives the batches (the C ABI releases the GIL)."""
        import pandas as pd
        if n_jobs < 2:
            warnings.warn("Processing in serial.", RuntimeWarning)
            return self.evaluate(df_data, atoms_key=atoms_key, energy_key=energy_key)
        order = np.arange(len(df_data))
        if shuffle:
            np.random.shuffle(order)
        batches = [df_data.iloc[part] for part in np.array_split(order, n_jobs) if len(part)]
        n_gpus = max(self._visible_devices(), 1)
        workers = [self._on_device(k % n_gpus) for k in range(len(batches))]
        kwargs = dict(atoms_key=atoms_key, energy_key=energy_key, progress=False)
        own_pool = None
        if client is None:
            from concurrent.futures import ThreadPoolExecutor
            client = own_pool = ThreadPoolExecutor(max_workers=min(n_gpus, len(batches)))
        try:
            futures = [client.submit(_evaluate_batch, worker, batch, kwargs)
                       for worker, batch in zip(workers, batches)]
            frames = [future.result() for future in futures]
        finally:
            if own_pool is not None:
                own_pool.shutdown()
        df_features = pd.concat(frames)
        names = df_features.index.get_level_values(0)
        # rows of one configuration stay together, configurations in the order of df_data
        rank = {name: k for k, name in enumerate(df_data.index)}
        position = np.argsort(np.array([rank[name] for name in names]), kind="stable")
        return df_features.iloc[position]

    def get_training_tuples(self, df_features, kappa, data_coordinator):
        """Deprecated in the reference as well (process.py:508-535)."""
        warnings.warn("get_training_tuples() is deprecated.", DeprecationWarning)
        return dataframe_to_training_tuples(df_features, kappa=kappa, energy_key=data_coordinator.energy_key)

    def batched_to_hdf(self, filename, df_data, client=None, n_jobs=16, batch_size=50, progress="bar",
                       table_template="features_{}", **kwargs):
        """Stream the feature rows of `df_data` into a chunked store, `batch_size` configurations per
        table (process.py:256-291).  Tables that are already in `filename` are skipped, so an
        interrupted run resumes where it stopped.  The store is the reference's HDF5 layout when
        PyTables is present and a chunk archive otherwise (`uf3_b200.store`); either is what
        `WeightedLinearModel.fit_from_file` reads."""
        import os
        from uf3_b200 import store
        idx_all = np.arange(len(df_data))
        idx_batches = np.array_split(idx_all, idx_all[batch_size::batch_size])
        digits = max(int(np.ceil(np.log10(len(idx_batches)) + 0.1)), 3)
        if os.path.isfile(filename):
            n_chunks, _, chunk_names, _ = store.analyze_hdf_tables(filename)
            warnings.warn(f"File already exists: contains {n_chunks} chunks.", RuntimeWarning)
        else:
            chunk_names = []
        kwargs["progress"] = False
        kwargs["n_jobs"] = n_jobs
        for j, idx_batch in enumerate(idx_batches):
            table_name = table_template.format(str(j).rjust(digits, "0"))
            if table_name in chunk_names:
                continue
            df_features = self.evaluate_parallel(df_data.iloc[idx_batch], client, **kwargs)
            store.save_feature_db(df_features, filename, table_name=table_name)


def _evaluate_batch(featurizer, batch, kwargs):
    """Module-level so that a process pool can pickle the call (util/parallel.py:182)."""
    return featurizer.evaluate(batch, **kwargs)


def save_feature_db(dataframe, filename, table_name="features"):
    """Reference name (process.py:538-548); the container is chosen by `uf3_b200.store`."""
    from uf3_b200 import store
    store.save_feature_db(dataframe, filename, table_name=table_name)


def load_feature_db(filename, table_name="features"):
    """Reference name (process.py:550-562)."""
    from uf3_b200 import store
    return store.load_feature_db(filename, table_name=table_name)


def dataframe_to_training_tuples(df_features, kappa=0.5, energy_key="energy"):
    """(x, y, w) of a feature DataFrame for a weighted fit (process.py:574-616): the first column is the
    target; energy rows weigh kappa / (n_e std_e) each, force rows (1 - kappa) / (n_f std_f)."""
    if kappa < 0 or kappa > 1:
        raise ValueError("Invalid domain for kappa weighting parameter.")
    if len(df_features) <= 1:
        raise ValueError(f"Not enough samples ({len(df_features)} provided)")
    is_energy = np.asarray(df_features.index.get_level_values(-1) == energy_key)
    data = df_features.to_numpy()
    y, x = data[:, 0], data[:, 1:]
    n_e, n_f = int(is_energy.sum()), int((~is_energy).sum())
    w = np.zeros(len(y))
    w[is_energy] = kappa / np.std(y[is_energy]) / n_e
    w[~is_energy] = (1 - kappa) / np.std(y[~is_energy]) / n_f
    return x, y, w


def flatten_by_interactions(vector_map, pair_tuples):
    return np.concatenate([vector_map[pair] for pair in pair_tuples], axis=-1)
