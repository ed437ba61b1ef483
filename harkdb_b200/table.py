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
* a DataFrame / csv / Arrow table whose columns differ in dtype keeps ONE DTYPE PER COLUMN on the device
  (``hark_table_from_columns``): integer keys stay integers next to float measures, which the reference's
  single homogeneous array cannot hold (README.md:8 blames the Futhark compiler).  ``get_data()`` is still the
  homogeneous 2-D array the reference would have built (``df.to_numpy()``, table.py:9).
* ``pyarrow.Table`` objects and ``.parquet`` files load column by column.
* a ``dict`` of 1-D arrays (name -> column) is taken as it is, WITHOUT a copy: the columns may live in pinned host memory
  (``hark_host_alloc``), and a query over such a non-resident table uploads only the columns it names.
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
    if file_name.endswith(".parquet"):
        import pyarrow.parquet as pq
        return load_arrow(pq.read_table(file_name))
    raise Exception("We do not support loading this file type")


def load_arrow(tbl):
    """pyarrow.Table -> (homogeneous values, headers), like load_df on the equivalent DataFrame."""
    cols = arrow_columns(tbl)
    values = np.column_stack(cols) if cols else np.zeros((tbl.num_rows, 0))
    return values, [str(nm).strip() for nm in tbl.column_names]


def arrow_columns(tbl):
    cols = []
    for i in range(tbl.num_columns):
        c = tbl.column(i)
        if c.null_count:
            raise Exception(f"column {tbl.column_names[i]} holds nulls; tables are dense numeric arrays")
        cols.append(np.ascontiguousarray(c.to_numpy()))
    return cols


def _is_arrow(obj):
    return type(obj).__module__.startswith("pyarrow") and hasattr(obj, "column_names") and hasattr(obj, "num_rows")


def source_columns(table):
    """Per-column host arrays of a source that has them (DataFrame, csv, Arrow, parquet), else None."""
    if isinstance(table, dict):
        return [np.asarray(c) for c in table.values()]
    if isinstance(table, pd.DataFrame):
        return [np.ascontiguousarray(table[c].to_numpy()) for c in table.columns]
    if _is_arrow(table):
        return arrow_columns(table)
    if isinstance(table, str) and table[-3:] == "csv":
        df = pd.read_csv(table)
        return [np.ascontiguousarray(df[c].to_numpy()) for c in df.columns]
    if isinstance(table, str) and table.endswith(".parquet"):
        import pyarrow.parquet as pq
        return arrow_columns(pq.read_table(table))
    return None


class _LazyStack:
    """Stand-in for the homogeneous 2-D array of a dict-of-columns table: built only if somebody asks for it."""

    def __init__(self, cols):
        self._cols = cols
        self.ndim = 2
        self.shape = (len(cols[0]) if cols else 0, len(cols))

    def __array__(self, dtype=None, copy=None):
        a = np.column_stack(self._cols) if self._cols else np.zeros((0, 0))
        return a.astype(dtype) if dtype is not None else a


def load_table(table_name, table):
    if isinstance(table, dict):
        cols = [np.asarray(c) for c in table.values()]
        if any(c.ndim != 1 or len(c) != len(cols[0]) for c in cols):
            raise Exception("a dict table needs 1-D columns of one length")
        return _LazyStack(cols), list(table.keys())
    if isinstance(table, pd.DataFrame):
        return load_df(table)
    elif isinstance(table, np.ndarray):
        return load_np(table)
    elif isinstance(table, str):
        return load_file(table)
    elif _is_arrow(table):
        return load_arrow(table)
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


class HostColumns(list):
    """Handle of a table that is NOT resident and whose columns differ in dtype: its per-column host arrays (already
    narrowed to their device dtypes).  FutharkContext uploads them for the one query, like the reference uploads its
    homogeneous array on every query."""


class Table:
    """A schema (list of column names) and a 2-D homogeneous array, optionally resident on the GPU."""

    def __init__(self, table_name, file_name):
        self._table_name = table_name
        table, headers = load_table(table_name, file_name)
        if not isinstance(table, _LazyStack):
            table = np.asarray(table)
        if table.ndim != 2:
            raise Exception("Table data must be two-dimensional")
        self._schema = list(headers)
        self._data = table
        self._device = None
        # one dtype per column when the source has them and they differ (else the homogeneous array is the table)
        self._columns = None
        cols = source_columns(file_name)
        if cols is not None and len(cols) == table.shape[1] and (len({c.dtype for c in cols}) > 1 or isinstance(file_name, dict)):
            for c, h in zip(cols, self._schema):
                if c.dtype.kind not in "iubf":
                    raise Exception(f"column {h} has dtype {c.dtype}; tables are numeric")
            self._columns = cols

    def get_schema(self):
        return self._schema

    def get_data(self):
        if isinstance(self._data, _LazyStack):
            self._data = np.asarray(self._data)
        return self._data

    def get_name(self):
        return self._table_name

    def is_u32_exact(self):
        """True when every column is an integer column whose values all lie in [0, 2^32): only then do the
        reference-pinned u32 entries (groupby.fut / join.fut compare and combine as u32) compute what SQL means; a
        table with negative values must take the typed entries (signed order, signed MIN / MAX)."""
        if getattr(self, "_u32_exact", None) is None:
            srcs = self._columns if self._columns is not None else [self._data]
            ok = True
            for a in srcs:
                a = np.asarray(a)
                if a.dtype.kind not in "iub":
                    ok = False
                elif a.size and (int(a.min()) < 0 or int(a.max()) >= 2 ** 32):
                    ok = False
            self._u32_exact = ok
        return self._u32_exact

    def _tag(self, handle):
        try:
            handle.u32_exact = self.is_u32_exact()
        except AttributeError:        # a plain ndarray cannot carry attributes: FutharkContext looks at its values
            pass
        return handle

    # ---- device residency (new) ----
    def upload(self, env):
        """Transpose to device SoA once; later queries use the resident handle."""
        if self._device is None:
            if self._columns is not None:
                self._device = env.from_columns([c.astype(entry_dtype(c), copy=False) for c in self._columns])
            else:
                self._device = env.to_device(self._data, entry_dtype(self._data))
            self._tag(self._device)
        return self._device

    def get_column_dtypes(self):
        """Device dtype of every column (what ``upload`` stores)."""
        if self._columns is not None:
            return [entry_dtype(c) for c in self._columns]
        return [entry_dtype(self._data)] * self._data.shape[1]

    def get_handle(self):
        """Resident device table if uploaded, else the host data (uploaded per query like the reference): the
        homogeneous array, or the per-column arrays when the columns differ in dtype."""
        if self._device is not None:
            return self._device
        if self._columns is not None:
            return self._tag(HostColumns(c.astype(entry_dtype(c), copy=False) for c in self._columns))
        return self._data

    def release(self):
        if self._device is not None:
            self._device.free()
            self._device = None
