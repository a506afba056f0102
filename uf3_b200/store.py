"""Chunked, resumable feature store — the hand-off between `BasisFeaturizer.batched_to_hdf`
and `WeightedLinearModel.fit_from_file`.

Reference: `/root/reference/uf3/representation/process.py:256-291` (`batched_to_hdf`: the data
frame is cut into batches of `batch_size` configurations, batch j becomes table
`features_{j:03d}`, tables that already exist are SKIPPED so an interrupted run resumes),
`process.py:538-562` (`save_feature_db` / `load_feature_db`), `data/io.py:943-970`
(`analyze_hdf_tables`, `dataframe_batch_loader`) and `regression/least_squares.py:355-433`
(`fit_from_file` walks the tables in sorted order).

Two containers behind the same four functions, chosen per file:
  * HDF5 through pandas / PyTables — the reference's own layout (`DataFrame.to_hdf(...,
    format='fixed')`), used when PyTables is importable and for any existing file that is HDF5;
  * a chunk archive (ZIP, stored uncompressed) for hosts without PyTables: table `name` is the
    three members `name/values.npy` (float64 rows, target in column 0), `name/index.json`
    (the (configuration, 'energy' | 'fx_i') pairs) and `name/columns.json`.  Appending a table
    rewrites only the archive's directory, so the resume semantics are the reference's.
Both hold exactly what `evaluate` returns, so a store written by either loads into the same
DataFrame.
"""
import io
import json
import os
import zipfile

import numpy as np

_HDF_MAGIC = b"\x89HDF\r\n\x1a\n"


def have_pytables():
    try:
        import tables  # noqa: F401
        return True
    except Exception:
        return False


def _is_hdf(filename):
    with open(filename, "rb") as fh:
        return fh.read(8) == _HDF_MAGIC


def _use_hdf(filename):
    if os.path.isfile(filename) and os.path.getsize(filename) > 0:
        return _is_hdf(filename)
    return have_pytables()


def _jsonable(value):
    if isinstance(value, (np.integer,)):
        return int(value)
    if isinstance(value, (np.floating,)):
        return float(value)
    if isinstance(value, (str, int, float)) or value is None:
        return value
    raise TypeError(f"index label {value!r} cannot be stored (str / int / float labels only)")


def save_feature_db(dataframe, filename, table_name="features"):
    """Append one table (process.py:538-548).  An existing table of that name is an error in the
    archive container (HDF5 'fixed' format silently replaces it; `batched_to_hdf` never does)."""
    if _use_hdf(filename):
        dataframe.to_hdf(filename, key=table_name, mode="a", format="fixed")
        return
    index = [[_jsonable(part) for part in (key if isinstance(key, tuple) else (key,))]
             for key in dataframe.index]
    values = io.BytesIO()
    np.save(values, np.ascontiguousarray(dataframe.to_numpy(dtype=np.float64)))
    with zipfile.ZipFile(filename, mode="a", compression=zipfile.ZIP_STORED, allowZip64=True) as archive:
        if f"{table_name}/values.npy" in archive.namelist():
            raise ValueError(f"table {table_name} already exists in {filename}")
        archive.writestr(f"{table_name}/index.json", json.dumps(index))
        archive.writestr(f"{table_name}/columns.json", json.dumps([str(c) for c in dataframe.columns]))
        archive.writestr(f"{table_name}/values.npy", values.getvalue())     # last: marks the table complete


def load_feature_db(filename, table_name="features"):
    """One table as the DataFrame `evaluate` returned (process.py:551-562)."""
    import pandas as pd
    if _is_hdf(filename):
        return pd.read_hdf(filename, table_name)
    with zipfile.ZipFile(filename, mode="r") as archive:
        index = json.loads(archive.read(f"{table_name}/index.json"))
        columns = json.loads(archive.read(f"{table_name}/columns.json"))
        values = np.load(io.BytesIO(archive.read(f"{table_name}/values.npy")))
    if index and len(index[0]) > 1:
        idx = pd.MultiIndex.from_tuples([tuple(key) for key in index])
    else:
        idx = pd.Index([key[0] for key in index])
    return pd.DataFrame(values, index=idx, columns=columns)


def analyze_hdf_tables(filename):
    """(n_chunks, n_entries, sorted chunk names, {name: rows}) of a store (data/io.py:943-956)."""
    if _is_hdf(filename):
        import tables
        lengths = {}
        with tables.open_file(filename, mode="r") as h5file:
            for group in h5file.list_nodes("/"):
                lengths[group._v_name] = h5file.get_node("/" + group._v_name, "axis0").nrows
    else:
        lengths = {}
        with zipfile.ZipFile(filename, mode="r") as archive:
            for member in archive.namelist():
                if member.endswith("/values.npy"):
                    with archive.open(member) as fh:
                        version = np.lib.format.read_magic(fh)
                        shape = (np.lib.format.read_array_header_1_0(fh) if version == (1, 0)
                                 else np.lib.format.read_array_header_2_0(fh))[0]
                    lengths[member[:-len("/values.npy")]] = int(shape[0])
    names = sorted(lengths)
    return len(names), int(sum(lengths.values())), names, lengths


def dataframe_batch_loader(filename, table_names):
    """Iterator over the tables of a store (data/io.py:959-970)."""
    for table_name in table_names:
        yield load_feature_db(filename, table_name)
