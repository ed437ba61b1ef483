"""Homogeneous numeric tables — same surface as the reference's table.py, plus a device-resident handle.

Mirrors /root/reference/table.py: ``load_df`` (:8-10), ``load_np`` (:12-16), ``load_file`` (:18-40),
``load_table`` (:42-50), ``Table`` (:52-80) with ``get_schema`` / ``get_data`` / ``get_name``; the
error messages are the reference's.  Differences, all invisible in query results:

* ``load_np`` names columns after the column count; the reference uses ``shape[0]`` (the row count,
  table.py:14), which leaves columns unnamed whenever a table has fewer rows than columns.
* ``Table.upload(env)`` transposes the row-major host array to device-resident SoA columns ONCE; the
  reference re-copies the whole array across the FFI on every query (FutharkContext.py:65,70).
* integer data is narrowed to the dtype the reference's entries take (i32, else u32, else i64) with a
  range check, because pandas hands the reference int64 (table.py:28) while main.fut's entries are
  i32/u32 (SURVEY.md §8b "dtype hazard").
"""

import numpy as np
import pandas as pd


def load_df(df):
    return df.to_numpy(), list(df)


def load_np(nparray, col_names=None):
    if col_names is None:
        return nparray, ["col" + str(i + 1) for i in range(nparray.shape[1])]
    return nparray, col_names


def load_file(file_name, col_names=None):
    """csv -> (values, headers) via pandas; txt -> np.loadtxt with c1..cn headers (table.py:18-40)."""
    if file_name[-3:] == "csv":
        table = pd.read_csv(file_name)
        headers = [h.strip() for h in table.columns.tolist()]
        return (table.values, headers)
    if file_name[-3:] == "txt":
        table = np.loadtxt(file_name, ndmin=2)
        if col_names is None:
            (_, dim_x) = table.shape
            headers = ["c" + str(i + 1) for i in range(dim_x)]
        else:
            headers = col_names
        return (table, headers)
    raise Exception("We do not support loading this file type")


def load_table(table_name, table):
    if isinstance(table, pd.DataFrame):
        return load_df(table)
    elif isinstance(table, np.ndarray):
        return load_np(table)
    elif isinstance(table, str):
        return load_file(table)
    else:
        raise Exception("Table is not in a file, numpy array or dataframe")


def entry_dtype(data: np.ndarray) -> np.dtype:
    """The device dtype a host array is stored as (see module docstring)."""
    if data.dtype.kind == "f":
        return np.dtype(np.float32) if data.dtype == np.float32 else np.dtype(np.float64)
    if data.dtype.kind in "iub":
        if data.size == 0:
            return np.dtype(np.int32)
        lo, hi = int(data.min()), int(data.max())
        if -(2 ** 31) <= lo and hi < 2 ** 31:
            return np.dtype(np.int32)
        if 0 <= lo and hi < 2 ** 32:
            return np.dtype(np.uint32)
        return np.dtype(np.int64)
    raise Exception(f"Table dtype {data.dtype} is not numeric")


class Table:
    """A schema (list of column names) and a 2-D homogeneous array, optionally resident on the GPU."""

    def __init__(self, table_name, file_name):
        self._table_name = table_name
        table, headers = load_table(table_name, file_name)
        table = np.asarray(table)
        if table.ndim != 2:
            raise Exception("Table data must be two-dimensional")
        self._schema = list(headers)
        self._data = table
        self._device = None

    def get_schema(self):
        return self._schema

    def get_data(self):
        return self._data

    def get_name(self):
        return self._table_name

    # ---- device residency (new) ----
    def upload(self, env):
        """Transpose to device SoA once; later queries use the resident handle."""
        if self._device is None:
            self._device = env.to_device(self._data, entry_dtype(self._data))
        return self._device

    def get_handle(self):
        """Resident device table if uploaded, else the host array (uploaded per query like the reference)."""
        return self._device if self._device is not None else self._data

    def release(self):
        if self._device is not None:
            self._device.free()
            self._device = None
