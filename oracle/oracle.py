"""ctypes front end of the CPU oracle (oracle/strsim_oracle.c) + a pure-Python twin.

TEST INFRASTRUCTURE ONLY -- see the header of strsim_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the product
package (polars-strsim_b200/).

The C functions restate /root/reference/src/expressions/strsim.rs:125-345; the pure-Python twin
below (`py_*`) is an independent second derivation (full-matrix DP, collections.Counter, explicit
flag lists) used to cross-check the C restatement on inputs the reference's golden vectors do not
cover (non-ASCII, long strings).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from collections import Counter
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libstrsim_oracle.so"

MEASURES = ("levenshtein", "jaro", "jaro_winkler", "jaccard", "sorensen_dice")
MEASURE_ID = {m: i for i, m in enumerate(MEASURES)}
N_INTS = 6


def build(force: bool = False) -> Path:
    """Compile the C oracle in place (gcc only; seconds)."""
    src = HERE / "strsim_oracle.c"
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "-B", "libstrsim_oracle.so"], check=True,
                       capture_output=True)
    return LIB_PATH


class _Column(ctypes.Structure):
    _fields_ = [
        ("views", ctypes.c_void_p),
        ("bufs", ctypes.POINTER(ctypes.c_void_p)),
        ("validity", ctypes.c_void_p),
        ("validity_offset", ctypes.c_int64),
        ("length", ctypes.c_int64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(LIB_PATH))
        L.oracle_pair.restype = ctypes.c_double
        L.oracle_pair.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p,
                                  ctypes.c_size_t, ctypes.c_void_p]
        L.oracle_batch_offsets.restype = None
        L.oracle_batch_offsets.argtypes = [ctypes.c_int, ctypes.c_int64] + [ctypes.c_void_p] * 9
        L.oracle_batch_views.restype = ctypes.c_int
        L.oracle_batch_views.argtypes = [ctypes.c_int, ctypes.POINTER(_Column),
                                         ctypes.POINTER(_Column), ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_int]
        _lib = L
    return _lib


def _mid(measure) -> int:
    return measure if isinstance(measure, int) else MEASURE_ID[measure]


def pair(measure, a: str | bytes, b: str | bytes):
    """(value, ints[6]) for one pair."""
    ba = a.encode("utf-8") if isinstance(a, str) else a
    bb = b.encode("utf-8") if isinstance(b, str) else b
    ints = np.zeros(N_INTS, dtype=np.int64)
    v = lib().oracle_pair(_mid(measure), ba, len(ba), bb, len(bb), ints.ctypes.data)
    return v, ints


def _pack(strings):
    n = len(strings)
    valid = np.ones(n, dtype=np.uint8)
    enc = []
    for i, s in enumerate(strings):
        if s is None:
            valid[i] = 0
            enc.append(b"")
        else:
            enc.append(s.encode("utf-8") if isinstance(s, str) else bytes(s))
    off = np.zeros(n + 1, dtype=np.int64)
    if n:
        np.cumsum([len(e) for e in enc], out=off[1:])
    data = np.frombuffer(b"".join(enc) + b"\0", dtype=np.uint8).copy()
    return data, off, valid


def batch(measure, a, b):
    """Row-wise oracle over two equal-length lists of str | bytes | None.

    Returns (values float64[n], valid bool[n], ints int64[n,6])."""
    assert len(a) == len(b)
    n = len(a)
    ad, ao, av = _pack(a)
    bd, bo, bv = _pack(b)
    out = np.zeros(n, dtype=np.float64)
    ov = np.zeros(n, dtype=np.uint8)
    ints = np.zeros((n, N_INTS), dtype=np.int64)
    lib().oracle_batch_offsets(_mid(measure), n, ad.ctypes.data, ao.ctypes.data, av.ctypes.data,
                               bd.ctypes.data, bo.ctypes.data, bv.ctypes.data, out.ctypes.data,
                               ov.ctypes.data, ints.ctypes.data)
    return out, ov.astype(bool), ints


def _column_from_arrow(arr, keep):
    """pyarrow string_view Array (one chunk) -> _Column (views pointer already offset)."""
    import pyarrow as pa

    assert arr.type == pa.string_view() or arr.type == pa.binary_view(), arr.type
    bufs = arr.buffers()
    validity, views, data = bufs[0], bufs[1], bufs[2:]
    ptrs = (ctypes.c_void_p * max(1, len(data)))(*[d.address if d is not None else None for d in data])
    keep.extend([bufs, ptrs])
    c = _Column()
    c.views = views.address + 16 * arr.offset
    c.bufs = ctypes.cast(ptrs, ctypes.POINTER(ctypes.c_void_p))
    c.validity = validity.address if validity is not None else None
    c.validity_offset = arr.offset
    c.length = len(arr)
    return c


def batch_views(measure, a, b, n_threads: int = 1, want_ints: bool = False):
    """Oracle over two single-chunk pyarrow string_view arrays, reading the Arrow buffers in place
    with the reference's static row-range partition over `n_threads` threads
    (/root/reference/src/expressions/strsim.rs:21-39,72-100)."""
    keep = []
    ca, cb = _column_from_arrow(a, keep), _column_from_arrow(b, keep)
    n = max(len(a), len(b)) if (len(a) != 1 or len(b) != 1) else 1
    out = np.zeros(n, dtype=np.float64)
    ov = np.zeros(n, dtype=np.uint8)
    ints = np.zeros((n, N_INTS), dtype=np.int64) if want_ints else None
    rc = lib().oracle_batch_views(_mid(measure), ctypes.byref(ca), ctypes.byref(cb), out.ctypes.data,
                                  ov.ctypes.data, ints.ctypes.data if want_ints else None, n_threads)
    if rc != 0:
        raise ValueError("Inputs must have the same length, or one of them must be a Utf8 literal.")
    return (out, ov.astype(bool), ints) if want_ints else (out, ov.astype(bool))


# --------------------------------------------------------------------------------------------------
# Pure-Python twin: independent derivation from the behavioural spec.
# --------------------------------------------------------------------------------------------------

def py_levenshtein(a: str, b: str):
    if a == b:
        return 1.0, None
    la, lb = len(a), len(b)
    D = [[0] * (lb + 1) for _ in range(la + 1)]
    for i in range(la + 1):
        D[i][0] = i
    for j in range(lb + 1):
        D[0][j] = j
    for i in range(1, la + 1):
        for j in range(1, lb + 1):
            D[i][j] = min(D[i - 1][j - 1] + (a[i - 1] != b[j - 1]), D[i - 1][j] + 1, D[i][j - 1] + 1)
    d = D[la][lb]
    return 1.0 - (float(d) / float(max(la, lb))), d


def py_jaro_mt(a: str, b: str):
    la, lb = len(a), len(b)
    bound = max(la, lb) // 2 - 1
    fa, fb = [False] * la, [False] * lb
    m = 0
    for i in range(min(la, lb + bound)):
        lo = max(0, i - bound)
        hi = min(i + bound, lb - 1)
        for j in range(lo, hi + 1):
            if a[i] == b[j] and not fb[j]:
                fa[i] = fb[j] = True
                m += 1
                break
    xs = [a[i] for i in range(la) if fa[i]]
    ys = [b[j] for j in range(lb) if fb[j]]
    t = sum(1 for x, y in zip(xs, ys) if x != y)
    return m, t


def py_jaro(a: str, b: str):
    if a == b:
        return 1.0
    if not a or not b:
        return 0.0
    if len(a) == 1 and len(b) == 1:
        return 1.0 if a == b else 0.0
    m, t = py_jaro_mt(a, b)
    if m == 0:
        return 0.0
    return (float(m) / float(len(a)) + float(m) / float(len(b)) + float(m - t // 2) / float(m)) / 3.0


def py_jaro_winkler(a: str, b: str):
    js = py_jaro(a, b)
    if js > 0.7:
        l = 0
        for x, y in zip(a[:4], b[:4]):
            if x != y:
                break
            l += 1
        return js + (float(l) * 0.1 * (1.0 - js))
    return js


def py_jaccard(a: str, b: str):
    if a == b:
        return 1.0
    if not a or not b:
        return 0.0
    ca, cb = Counter(a), Counter(b)
    inter = sum((ca & cb).values())
    uni = sum((ca | cb).values())
    return float(inter) / float(uni)


def py_sorensen_dice(a: str, b: str):
    if a == b:
        return 1.0
    if not a or not b:
        return 0.0
    ca, cb = Counter(a), Counter(b)
    inter = sum((ca & cb).values())
    return 2.0 * float(inter) / float(len(a) + len(b))


PY_TWIN = {
    "levenshtein": lambda a, b: py_levenshtein(a, b)[0],
    "jaro": py_jaro,
    "jaro_winkler": py_jaro_winkler,
    "jaccard": py_jaccard,
    "sorensen_dice": py_sorensen_dice,
}


def n_host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
