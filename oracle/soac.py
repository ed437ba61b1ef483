"""Futhark SOAC semantics in pure Python — TEST INFRASTRUCTURE ONLY.

This file is part of the parity oracle.  Nothing under ``harkdb_b200/`` may import it;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may.

The reference (philass/HarkDB) is written in Futhark, whose compiler is not available in
this image.  To pin the operator semantics as tightly as possible without
it, the second-order array combinators the reference's sources use are restated here with
Futhark's *sequential* meaning (what ``futhark c``, the backend ``setup.sh:12`` builds,
executes), and ``oracle/hark_ref.py`` re-expresses the reference programs line by line on
top of them.  Everything is plain Python lists and left-to-right loops: small cases only.

Semantics followed (Futhark language reference, sequential backend):
  scan op ne xs      inclusive prefix, acc starts at ne, acc = op(acc, x) left to right
  reduce op ne xs    left fold from ne
  scatter dst is vs  dst[is[j]] = vs[j] for in-bounds is[j]; out-of-bounds writes ignored
  rotate r xs        result[i] = xs[(i + r) mod n]
  partition p xs     (stable kept, stable rest)
  reduce_by_index dst op ne is vs   dst[is[j]] = op(dst[is[j]], vs[j]) for in-bounds is[j]
"""

from __future__ import annotations

from typing import Callable, List, Sequence, Tuple, TypeVar

T = TypeVar("T")
U = TypeVar("U")

U32 = 0xFFFFFFFF


def u32(x: int) -> int:
    return x & U32


def i32(x: int) -> int:
    x &= U32
    return x - (1 << 32) if x & 0x80000000 else x


def iota(n: int) -> List[int]:
    return list(range(n))


def replicate(n: int, v: T) -> List[T]:
    return [v for _ in range(n)]


def fmap(f: Callable[[T], U], xs: Sequence[T]) -> List[U]:
    return [f(x) for x in xs]


def map2(f: Callable, xs: Sequence, ys: Sequence) -> List:
    assert len(xs) == len(ys), "map2: size mismatch is a Futhark type error"
    return [f(x, y) for x, y in zip(xs, ys)]


def scan(op: Callable[[T, T], T], ne: T, xs: Sequence[T]) -> List[T]:
    out: List[T] = []
    acc = ne
    for x in xs:
        acc = op(acc, x)
        out.append(acc)
    return out


def reduce(op: Callable[[T, T], T], ne: T, xs: Sequence[T]) -> T:
    acc = ne
    for x in xs:
        acc = op(acc, x)
    return acc


def scatter(dst: Sequence[T], idxs: Sequence[int], vals: Sequence[T]) -> List[T]:
    out = list(dst)
    n = len(out)
    assert len(idxs) == len(vals)
    for i, v in zip(idxs, vals):
        if 0 <= i < n:
            out[i] = v
    return out


def rotate(r: int, xs: Sequence[T]) -> List[T]:
    n = len(xs)
    if n == 0:
        return []
    return [xs[(i + r) % n] for i in range(n)]


def partition(p: Callable[[T], bool], xs: Sequence[T]) -> Tuple[List[T], List[T]]:
    return [x for x in xs if p(x)], [x for x in xs if not p(x)]


def reduce_by_index(dst: Sequence[T], op: Callable[[T, T], T], ne: T,
                    idxs: Sequence[int], vals: Sequence[T]) -> List[T]:
    out = list(dst)
    n = len(out)
    for i, v in zip(idxs, vals):
        if 0 <= i < n:
            out[i] = op(out[i], v)
    return out


def concat(*xss: Sequence[T]) -> List[T]:
    out: List[T] = []
    for xs in xss:
        out.extend(xs)
    return out


def last(xs: Sequence[T]) -> T:
    return xs[len(xs) - 1]
