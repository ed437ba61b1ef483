"""harkdb_b200 — B200-native operator path for HarkDB behind HarkDB's own Python API.

    from harkdb_b200 import FutharkContext
    fc = FutharkContext()
    fc.create_table('game_1', 'data.csv')
    fc.sql("select col1, col3 from game_1")

Importing this package never touches CUDA; constructing a FutharkContext does and raises when
libhark.so or a B200 is missing (there is no CPU fallback).
"""

from .FutharkContext import FutharkContext  # noqa: F401
from .table import Table  # noqa: F401
from .parse import sql_parse  # noqa: F401

__all__ = ["FutharkContext", "Table", "sql_parse"]
