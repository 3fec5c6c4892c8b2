"""ctypes binding of libpolars_strsim_b200.so (C ABI declared in include/strsim_b200.h)."""
from __future__ import annotations

import ctypes
from pathlib import Path

import numpy as np

LIB_NAME = "libpolars_strsim_b200.so"
import os as _os

LIB_PATH = Path(_os.environ.get("STRSIM_B200_LIB") or (Path(__file__).parent / LIB_NAME))

MEASURES = ("levenshtein", "jaro", "jaro_winkler", "jaccard", "sorensen_dice")
MEASURE_ID = {m: i for i, m in enumerate(MEASURES)}
DBG_INTS = 6
STATUS = {1: "ShapeMismatch", 2: "SchemaMismatch", 3: "InvalidArgument", 4: "CudaError", 5: "OutOfMemory"}


class StrsimError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS.get(status, status)}: {message}")
        self.status = status


class ViewChunk(ctypes.Structure):
    _fields_ = [
        ("views", ctypes.c_void_p),
        ("validity", ctypes.c_void_p),
        ("offset", ctypes.c_int64),
        ("length", ctypes.c_int64),
        ("data_buffers", ctypes.POINTER(ctypes.c_void_p)),
        ("data_buffer_sizes", ctypes.POINTER(ctypes.c_int64)),
        ("n_data_buffers", ctypes.c_int64),
    ]


class ArrowSchema(ctypes.Structure):
    pass


class ArrowArray(ctypes.Structure):
    pass


ArrowSchema._fields_ = [
    ("format", ctypes.c_char_p), ("name", ctypes.c_char_p), ("metadata", ctypes.c_char_p),
    ("flags", ctypes.c_int64), ("n_children", ctypes.c_int64),
    ("children", ctypes.POINTER(ctypes.POINTER(ArrowSchema))), ("dictionary", ctypes.POINTER(ArrowSchema)),
    ("release", ctypes.CFUNCTYPE(None, ctypes.POINTER(ArrowSchema))), ("private_data", ctypes.c_void_p),
]
ArrowArray._fields_ = [
    ("length", ctypes.c_int64), ("null_count", ctypes.c_int64), ("offset", ctypes.c_int64),
    ("n_buffers", ctypes.c_int64), ("n_children", ctypes.c_int64),
    ("buffers", ctypes.POINTER(ctypes.c_void_p)), ("children", ctypes.POINTER(ctypes.POINTER(ArrowArray))),
    ("dictionary", ctypes.POINTER(ArrowArray)),
    ("release", ctypes.CFUNCTYPE(None, ctypes.POINTER(ArrowArray))), ("private_data", ctypes.c_void_p),
]

_lib = None


def lib():
    """Load the CUDA library; there is deliberately no fallback when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C polars-strsim_b200/csrc` "
            "(or __graft_entry__.build()); polars-strsim_b200 has no CPU fallback"
        )
    L = ctypes.CDLL(str(LIB_PATH))
    P, I64, SZ = ctypes.c_void_p, ctypes.c_int64, ctypes.c_size_t
    L.strsim_b200_compute_host.restype = ctypes.c_int
    L.strsim_b200_compute_host.argtypes = [ctypes.c_int, ctypes.POINTER(ViewChunk), SZ,
                                           ctypes.POINTER(ViewChunk), SZ, P, P, ctypes.POINTER(I64), P]
    L.strsim_b200_compute_host_multi.restype = ctypes.c_int
    L.strsim_b200_compute_host_multi.argtypes = [ctypes.POINTER(ctypes.c_int), SZ, ctypes.POINTER(ViewChunk), SZ,
                                                 ctypes.POINTER(ViewChunk), SZ, ctypes.POINTER(P), P,
                                                 ctypes.POINTER(I64), ctypes.POINTER(P)]
    L.strsim_b200_compute_arrow.restype = ctypes.c_int
    L.strsim_b200_compute_arrow.argtypes = [ctypes.c_int, P, P, SZ, P, P, SZ, P]
    L.strsim_b200_column_upload.restype = ctypes.c_int
    L.strsim_b200_column_upload.argtypes = [ctypes.POINTER(ViewChunk), SZ, ctypes.POINTER(P)]
    L.strsim_b200_column_free.restype = None
    L.strsim_b200_column_free.argtypes = [P]
    L.strsim_b200_column_length.restype = I64
    L.strsim_b200_column_length.argtypes = [P]
    L.strsim_b200_column_algorithmic_bytes.restype = I64
    L.strsim_b200_column_algorithmic_bytes.argtypes = [P]
    L.strsim_b200_column_restat.restype = ctypes.c_int
    L.strsim_b200_column_restat.argtypes = [P, P]
    L.strsim_b200_compute_device.restype = ctypes.c_int
    L.strsim_b200_compute_device.argtypes = [ctypes.c_int, P, P, P, P, P, P]
    L.strsim_b200_compute_device_multi.restype = ctypes.c_int
    L.strsim_b200_compute_device_multi.argtypes = [ctypes.POINTER(ctypes.c_int), SZ, P, P, ctypes.POINTER(P), P,
                                                   ctypes.POINTER(P), P]
    L.strsim_b200_set_device.restype = ctypes.c_int
    L.strsim_b200_set_device.argtypes = [ctypes.c_int]
    L.strsim_b200_device_count.restype = ctypes.c_int
    L.strsim_b200_last_error.restype = ctypes.c_char_p
    L.strsim_b200_kernel_launches.restype = ctypes.c_uint64
    L.strsim_b200_last_overflow.restype = None
    L.strsim_b200_last_overflow.argtypes = [ctypes.POINTER(I64)]
    L.strsim_b200_last_redo_slices.restype = ctypes.c_int
    L.strsim_b200_version.restype = ctypes.c_char_p
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise StrsimError(rc, lib().strsim_b200_last_error().decode("utf-8", "replace"))


def measure_id(measure) -> int:
    return measure if isinstance(measure, int) else MEASURE_ID[measure]


def as_chunks(col):
    """pyarrow Array / ChunkedArray of string_view (or string / large_string, which pyarrow casts to
    views zero-copy) -> (ctypes array of ViewChunk, keepalive list).  Other types are a SchemaMismatch."""
    import pyarrow as pa

    if isinstance(col, (str, bytes)) or col is None:
        col = pa.array([col], type=pa.string_view())
    arrays = col.chunks if isinstance(col, pa.ChunkedArray) else [col]
    keep = []
    out = (ViewChunk * max(1, len(arrays)))()
    for i, arr in enumerate(arrays):
        if arr.type in (pa.string(), pa.large_string()):
            arr = arr.cast(pa.string_view())
        if arr.type != pa.string_view():
            # String columns only, like the reference's `inputs[i].str()?` (strsim.rs:46-47): binary data
            # carries no UTF-8 guarantee
            raise StrsimError(2, f"invalid series dtype: expected `String`, got `{arr.type}`")
        bufs = arr.buffers()
        data = [b for b in bufs[2:]]
        nd = len(data)
        ptrs = (ctypes.c_void_p * max(1, nd))(*[(d.address if d is not None else None) for d in data])
        sizes = (ctypes.c_int64 * max(1, nd))(*[(d.size if d is not None else 0) for d in data])
        keep += [arr, bufs, ptrs, sizes]
        ch = out[i]
        ch.views = bufs[1].address if bufs[1] is not None else None
        ch.validity = bufs[0].address if bufs[0] is not None else None
        ch.offset = arr.offset
        ch.length = len(arr)
        ch.data_buffers = ctypes.cast(ptrs, ctypes.POINTER(ctypes.c_void_p))
        ch.data_buffer_sizes = ctypes.cast(sizes, ctypes.POINTER(ctypes.c_int64))
        ch.n_data_buffers = nd
    return out, len(arrays), keep


def total_length(col) -> int:
    return 1 if isinstance(col, (str, bytes)) or col is None else len(col)


def compute_host(measure, a, b, debug: bool = False, out_values=None, out_validity=None, prepared=None):
    """End-to-end call with host buffers: returns (values f64[n], valid bool[n], null_count[, ints]).

    out_values / out_validity: optional preallocated numpy arrays (e.g. views of pinned memory) of
    n float64 / ceil(n/8)+8 uint8; with out_validity given, the packed bitmap is returned instead of
    a bool array.  prepared: result of prepare(a, b) to keep chunk marshalling out of a timed loop."""
    L = lib()
    ca, na, cb, nb, n, _keep = prepared if prepared is not None else prepare(a, b)
    values = out_values if out_values is not None else np.zeros(max(n, 1), dtype=np.float64)
    vbytes = out_validity if out_validity is not None else np.zeros((max(n, 1) + 7) // 8 + 8, dtype=np.uint8)
    ints = np.zeros((max(n, 1), DBG_INTS), dtype=np.int32) if debug else None
    nulls = ctypes.c_int64(0)
    rc = L.strsim_b200_compute_host(measure_id(measure), ca, na, cb, nb, values.ctypes.data, vbytes.ctypes.data,
                                    ctypes.byref(nulls), ints.ctypes.data if debug else None)
    _check(rc)
    if out_validity is not None:
        valid = vbytes
    else:
        valid = np.unpackbits(vbytes, bitorder="little")[:n].astype(bool)
    values = values[:n]
    if debug:
        return values, valid, nulls.value, ints[:n]
    return values, valid, nulls.value


def compute_host_multi(measures, a, b, out_values=None, out_validity=None, prepared=None, debug: bool = False):
    """Several measures over ONE upload of the two columns and (for distinct measures) ONE fused
    kernel pass (strsim_b200_compute_host_multi).

    Returns (list of float64 arrays, validity, null_count[, list of int32[n,6] records]); validity is a
    bool array unless a preallocated packed bitmap `out_validity` was passed."""
    L = lib()
    ca, na, cb, nb, n, _keep = prepared if prepared is not None else prepare(a, b)
    k = len(measures)
    ids = (ctypes.c_int * k)(*[measure_id(m) for m in measures])
    outs = out_values if out_values is not None else [np.zeros(max(n, 1), dtype=np.float64) for _ in range(k)]
    vbytes = out_validity if out_validity is not None else np.zeros((max(n, 1) + 7) // 8 + 8, dtype=np.uint8)
    ptrs = (ctypes.c_void_p * k)(*[o.ctypes.data for o in outs])
    ints = [np.zeros((max(n, 1), DBG_INTS), dtype=np.int32) for _ in range(k)] if debug else None
    iptrs = (ctypes.c_void_p * k)(*[i.ctypes.data for i in ints]) if debug else None
    nulls = ctypes.c_int64(0)
    _check(L.strsim_b200_compute_host_multi(ids, k, ca, na, cb, nb, ptrs, vbytes.ctypes.data, ctypes.byref(nulls), iptrs))
    valid = vbytes if out_validity is not None else np.unpackbits(vbytes, bitorder="little")[:n].astype(bool)
    if debug:
        return [o[:n] for o in outs], valid, nulls.value, [i[:n] for i in ints]
    return [o[:n] for o in outs], valid, nulls.value


def prepare(a, b):
    """Marshal two Arrow columns into ViewChunk arrays once (reusable across calls)."""
    ca, na, keep_a = as_chunks(a)
    cb, nb, keep_b = as_chunks(b)
    la, lb = total_length(a), total_length(b)
    n = lb if la == 1 else la
    if la != lb and la != 1 and lb != 1:
        n = 0  # the library reports the shape error
    return ca, na, cb, nb, n, (keep_a, keep_b)


class DeviceColumn:
    """A String column resident in HBM (views, validity and data buffers copied as they are)."""

    def __init__(self, col):
        L = lib()
        chunks, n, keep = as_chunks(col)
        h = ctypes.c_void_p()
        _check(L.strsim_b200_column_upload(chunks, n, ctypes.byref(h)))
        self._h = h

    @property
    def handle(self):
        return self._h

    def __len__(self):
        return int(lib().strsim_b200_column_length(self._h))

    @property
    def algorithmic_bytes(self) -> int:
        return int(lib().strsim_b200_column_algorithmic_bytes(self._h))

    def restat(self, stream: int = 0):
        """re-runs the column-statistics pre-pass (what an upload does once) on `stream` and waits for it"""
        _check(lib().strsim_b200_column_restat(self._h, stream or None))

    def free(self):
        if self._h:
            lib().strsim_b200_column_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def compute_device(measure, a: DeviceColumn, b: DeviceColumn, out_ptr: int, validity_ptr: int = 0,
                   dbg_ptr: int = 0, stream: int = 0):
    """Launch on device-resident columns; pointers are raw device addresses (e.g. tensor.data_ptr())."""
    _check(lib().strsim_b200_compute_device(measure_id(measure), a.handle, b.handle, out_ptr,
                                            validity_ptr or None, dbg_ptr or None, stream or None))


def compute_device_multi(measures, a: DeviceColumn, b: DeviceColumn, out_ptrs, validity_ptr: int = 0, dbg_ptrs=None,
                         stream: int = 0):
    """Several distinct measures over resident columns in ONE fused pass; out_ptrs[k] (raw device
    addresses of n float64) receives measures[k]."""
    k = len(measures)
    ids = (ctypes.c_int * k)(*[measure_id(m) for m in measures])
    outs = (ctypes.c_void_p * k)(*out_ptrs)
    dbgs = (ctypes.c_void_p * k)(*[(p or None) for p in dbg_ptrs]) if dbg_ptrs else None
    _check(lib().strsim_b200_compute_device_multi(ids, k, a.handle, b.handle, outs, validity_ptr or None, dbgs,
                                                  stream or None))


def set_device(device: int):
    _check(lib().strsim_b200_set_device(device))


def bind_thread_near_device(device: int) -> int:
    """Binds the calling thread (threads and pinned buffers created from now on follow) to the NUMA node
    of the device's PCIe link; -1 = nothing to do."""
    return int(lib().strsim_b200_bind_thread_near_device(device))


def kernel_launches() -> int:
    return int(lib().strsim_b200_kernel_launches())


def last_redo_slices() -> int:
    return int(lib().strsim_b200_last_redo_slices())


def last_overflow():
    out = (ctypes.c_int64 * 2)()
    lib().strsim_b200_last_overflow(out)
    return int(out[0]), int(out[1])
