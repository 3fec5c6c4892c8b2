"""The C ABI boundary without a GPU: every symbol include/strsim_b200.h declares is exported, the
planning-time plugin functions work, and the compute entry points fail LOUDLY (no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "strsim_b200.h"
MEASURES = ("levenshtein", "jaro", "jaro_winkler", "jaccard", "sorensen_dice")


@pytest.fixture(scope="module")
def native():
    from polars_strsim import _native

    return _native


def declared_symbols():
    text = HEADER.read_text()
    names = set(re.findall(r"STRSIM_API\s+[\w\s\*]+?\b(\w+)\s*\(", text))
    names -= {"void"}
    for m in re.findall(r"STRSIM_DECLARE_PLUGIN\((\w+)\)", text):
        if m != "name":
            names |= {f"_polars_plugin_{m}", f"_polars_plugin_field_{m}"}
    return {n for n in names if "##" not in n}


def test_every_declared_symbol_is_exported(native):
    L = native.lib()
    syms = declared_symbols()
    assert {"strsim_b200_compute_host", "strsim_b200_compute_host_multi", "strsim_b200_compute_arrow",
            "strsim_b200_column_upload", "strsim_b200_compute_device", "_polars_plugin_get_version",
            "_polars_plugin_get_last_error_message"} <= syms
    for m in MEASURES:
        assert f"_polars_plugin_{m}" in syms and f"_polars_plugin_field_{m}" in syms
    for s in sorted(syms):
        getattr(L, s)  # AttributeError if not exported


def test_plugin_version_and_field(native):
    import pyarrow as pa
    from polars_strsim._native import ArrowSchema

    L = native.lib()
    L._polars_plugin_get_version.restype = ctypes.c_uint32
    assert L._polars_plugin_get_version() == 1  # (major 0 << 16) + minor 1, polars-ffi version_0
    # field function: Float64 named like the first input (mod.rs:8 `output_type=Float64`)
    fields = (ArrowSchema * 2)()
    pa.field("name_a", pa.string_view())._export_to_c(ctypes.addressof(fields[0]))
    pa.field("name_b", pa.string_view())._export_to_c(ctypes.addressof(fields[1]))
    out = ArrowSchema()
    for m in MEASURES:
        getattr(L, f"_polars_plugin_field_{m}")(fields, ctypes.c_size_t(2), ctypes.byref(out))
        f = pa.Field._import_from_c(ctypes.addressof(out))
        assert f.name == "name_a" and f.type == pa.float64()
    for f in fields:
        f.release(ctypes.byref(f))


class SeriesExport(ctypes.Structure):
    pass


SeriesExport._fields_ = [("field", ctypes.c_void_p), ("arrays", ctypes.POINTER(ctypes.c_void_p)),
                         ("len", ctypes.c_size_t),
                         ("release", ctypes.CFUNCTYPE(None, ctypes.POINTER(SeriesExport))),
                         ("private_data", ctypes.c_void_p)]


def make_series(arr, released, keep):
    """Fabricates what polars-ffi's export_series hands a plugin: boxed schema + boxed arrays."""
    import pyarrow as pa
    from polars_strsim._native import ArrowArray, ArrowSchema

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

    cb = SeriesExport._fields_[3][1](_release)
    se = SeriesExport(ctypes.addressof(schema), ptrs, len(chunks), cb, 1)
    keep += [schema, c_arrays, ptrs, cb]
    return se, c_arrays


def test_plugin_call_fails_loudly_without_gpu(native):
    """No GPU here: the call must leave return_value untouched, consume its inputs and store a
    message -- never compute on the CPU."""
    import pyarrow as pa

    L = native.lib()
    if L.strsim_b200_device_count() > 0:
        pytest.skip("a CUDA device is present; the GPU variant lives in test_gpu_parity.py")
    released, keep = [], []
    inputs = (SeriesExport * 2)()
    a, arrs_a = make_series(pa.array(["phillips", None], type=pa.string_view()), released, keep)
    b, arrs_b = make_series(pa.array(["philips", "x"], type=pa.string_view()), released, keep)
    inputs[0], inputs[1] = a, b
    ret = SeriesExport()
    ctx = ctypes.c_uint64(0)
    L._polars_plugin_jaro_winkler(inputs, ctypes.c_size_t(2), None, ctypes.c_size_t(0), ctypes.byref(ret),
                                  ctypes.byref(ctx))
    assert not ret.private_data and not ret.release
    assert len(released) == 2                                   # both SeriesExports released
    assert all(not x.release for x in arrs_a + arrs_b)          # and every chunk's contents
    L._polars_plugin_get_last_error_message.restype = ctypes.c_char_p
    assert b"no CPU fallback" in L._polars_plugin_get_last_error_message()
    with pytest.raises(native.StrsimError, match="no CPU fallback"):
        native.compute_host("jaro", pa.array(["a"], type=pa.string_view()), pa.array(["b"], type=pa.string_view()))


def test_bind_thread_near_device_is_a_no_op_without_topology(native):
    """strsim_b200_bind_thread_near_device: -1 (nothing done, affinity untouched) when there is no device
    or the host exposes no NUMA locality for it -- never an error."""
    import os

    before = os.sched_getaffinity(0)
    node = native.bind_thread_near_device(0)
    assert node >= -1
    if node == -1:
        assert os.sched_getaffinity(0) == before
    else:
        assert os.sched_getaffinity(0) <= before
        os.sched_setaffinity(0, before)
