"""The five measures over pyarrow arrays, through the Arrow C Data Interface entry point
(`strsim_b200_compute_arrow`).  Same names and argument meaning as the Polars functions: two
String columns (or one column and a scalar literal) -> Float64 with nulls propagated.

    >>> import pyarrow as pa
    >>> from polars_strsim.arrow import jaro_winkler
    >>> jaro_winkler(pa.array(["phillips"]), pa.array(["philips"]))   # -> [0.975]
"""
from __future__ import annotations

import ctypes

from polars_strsim import _native
from polars_strsim._native import ArrowArray, ArrowSchema, StrsimError


def _export(col):
    import pyarrow as pa

    if isinstance(col, (str, bytes)) or col is None:
        col = pa.array([col], type=pa.string_view())
    arrays = col.chunks if isinstance(col, pa.ChunkedArray) else [col]
    if not arrays:
        arrays = [pa.array([], type=col.type)]
    c_arrays = [ArrowArray() for _ in arrays]
    c_schemas = [ArrowSchema() for _ in arrays]
    for arr, ca, cs in zip(arrays, c_arrays, c_schemas):
        arr._export_to_c(ctypes.addressof(ca), ctypes.addressof(cs))
    ptrs = (ctypes.c_void_p * len(arrays))(*[ctypes.addressof(a) for a in c_arrays])
    return c_arrays, c_schemas, ptrs


def _release(c_arrays, c_schemas):
    for a in c_arrays:
        if a.release:
            a.release(ctypes.byref(a))
    for s in c_schemas:
        if s.release:
            s.release(ctypes.byref(s))


def compute(measure, a, b):
    """-> pyarrow.DoubleArray (validity = AND of the inputs' validities)."""
    import pyarrow as pa

    L = _native.lib()
    aa, sa, pa_ptrs = _export(a)
    ab, sb, pb_ptrs = _export(b)
    out = ArrowArray()
    try:
        rc = L.strsim_b200_compute_arrow(_native.measure_id(measure), ctypes.addressof(sa[0]), pa_ptrs, len(aa),
                                         ctypes.addressof(sb[0]), pb_ptrs, len(ab), ctypes.addressof(out))
    finally:
        _release(aa, sa)
        _release(ab, sb)
    if rc != 0:
        raise StrsimError(rc, L.strsim_b200_last_error().decode("utf-8", "replace"))
    return pa.Array._import_from_c(ctypes.addressof(out), pa.float64())


def levenshtein(expr, other):
    return compute("levenshtein", expr, other)


def jaro(expr, other):
    return compute("jaro", expr, other)


def jaro_winkler(expr, other):
    return compute("jaro_winkler", expr, other)


def jaccard(expr, other):
    return compute("jaccard", expr, other)


def sorensen_dice(expr, other):
    return compute("sorensen_dice", expr, other)


__all__ = ["compute", "levenshtein", "jaro", "jaro_winkler", "jaccard", "sorensen_dice"]
