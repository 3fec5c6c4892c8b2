"""Drives the `_polars_plugin_<measure>` symbols the way Polars does (polars-ffi 0.43.1 `version_0`,
SURVEY.md 8(b)): the inputs are exported as boxed `SeriesExport`s that the callee consumes, the result
comes back as a `SeriesExport` holding one Float64 Arrow array in buffers the plugin allocated.

Bench / test support (Polars itself is not installed in this image): used by bench.py's `e2e` leg,
tools/plugin_e2e.py and the GPU tests.  The reference side of the same call is
`register_plugin_function(..., function_name=<measure>, args=[expr, other])`
(/root/reference/polars_strsim/__init__.py:11-16): one call per expression.
"""
from __future__ import annotations

import ctypes

import numpy as np
import pyarrow as pa

from polars_strsim import _native
from polars_strsim._native import ArrowArray, ArrowSchema

MEASURES = _native.MEASURES


class SeriesExport(ctypes.Structure):
    pass


_RELEASE = ctypes.CFUNCTYPE(None, ctypes.POINTER(SeriesExport))
SeriesExport._fields_ = [("field", ctypes.c_void_p), ("arrays", ctypes.POINTER(ctypes.c_void_p)),
                         ("len", ctypes.c_size_t), ("release", _RELEASE), ("private_data", ctypes.c_void_p)]


def make_series(arr, released, keep):
    """Fabricates what polars-ffi's export_series hands a plugin: boxed schema + boxed arrays."""
    chunks = arr.chunks if isinstance(arr, pa.ChunkedArray) else [arr]
    schema = ArrowSchema()
    pa.field("", chunks[0].type)._export_to_c(ctypes.addressof(schema))
    c_arrays = [ArrowArray() for _ in chunks]
    for ch, ca in zip(chunks, c_arrays):
        ch._export_to_c(ctypes.addressof(ca))
    ptrs = (ctypes.c_void_p * len(chunks))(*[ctypes.addressof(a) for a in c_arrays])

    def _release(p):
        released.append(1)
        if schema.release:
            schema.release(ctypes.byref(schema))
        p.contents.release = ctypes.cast(None, type(p.contents.release))

    cb = _RELEASE(_release)
    se = SeriesExport(ctypes.addressof(schema), ptrs, len(chunks), cb, 1)
    keep += [schema, c_arrays, ptrs, cb]
    return se, c_arrays


class PluginResult:
    """The Series a plugin call returned: one or more Float64 chunks living in the plugin's own buffers."""

    def __init__(self, ret: SeriesExport):
        self._ret = ret
        self.arrays = [ArrowArray.from_address(ret.arrays[i]) for i in range(ret.len)]

    def __len__(self):
        return sum(int(a.length) for a in self.arrays)

    def values(self):
        """zero-copy float64 views of the chunks' value buffers (valid until release())"""
        out = []
        for a in self.arrays:
            n = int(a.length)
            buf = (ctypes.c_double * max(n, 1)).from_address(a.buffers[1])
            out.append(np.frombuffer(buf, dtype=np.float64, count=n))
        return out

    def validity(self):
        """bool array over all rows"""
        out = []
        for a in self.arrays:
            n = int(a.length)
            if not a.buffers[0]:
                out.append(np.ones(n, dtype=bool))
                continue
            nb = (int(a.offset) + n + 7) // 8
            raw = np.frombuffer((ctypes.c_uint8 * max(nb, 1)).from_address(a.buffers[0]), dtype=np.uint8, count=nb)
            out.append(np.unpackbits(raw, bitorder="little")[int(a.offset):int(a.offset) + n].astype(bool))
        return np.concatenate(out) if out else np.zeros(0, dtype=bool)

    @property
    def null_count(self):
        return sum(int(a.null_count) for a in self.arrays)

    def release(self):
        """what polars-ffi's import does: take the arrays' contents, then release the SeriesExport"""
        for a in self.arrays:
            if a.release:
                a.release(ctypes.byref(a))
        if self._ret.release:
            self._ret.release(ctypes.byref(self._ret))
        self.arrays = []


def call(measure: str, a, b, context_flags: int = 0) -> PluginResult:
    """One `_polars_plugin_<measure>` call over two Arrow columns; raises with the plugin's message when
    `return_value` comes back untouched."""
    L = _native.lib()
    released, keep = [], []
    inputs = (SeriesExport * 2)()
    inputs[0], _ = make_series(a, released, keep)
    inputs[1], _ = make_series(b, released, keep)
    ret = SeriesExport()
    ctx = ctypes.c_uint64(context_flags)
    getattr(L, f"_polars_plugin_{measure}")(inputs, ctypes.c_size_t(2), None, ctypes.c_size_t(0), ctypes.byref(ret),
                                            ctypes.byref(ctx))
    if not ret.private_data:
        L._polars_plugin_get_last_error_message.restype = ctypes.c_char_p
        raise RuntimeError("the plugin failed with message: " +
                           L._polars_plugin_get_last_error_message().decode("utf-8", "replace"))
    return PluginResult(ret)


def cache_clear():
    L = _native.lib()
    L.strsim_b200_cache_clear.restype = None
    L.strsim_b200_cache_clear()
