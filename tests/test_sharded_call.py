"""One host / plugin call sharded over several GPUs inside the library (STRSIM_B200_DEVICES): the GPU
counterpart of the reference's fan-out over Polars' pool inside one call (strsim.rs:72-104: split_offsets,
one task per range, chunks re-assembled).  Needs at least two devices (`gpurun --gpus 2`); the
partitioning rule itself is covered on the CPU by tests/test_sharding.py."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

CODE = r"""
import ctypes, sys, random
import numpy as np
sys.path[:0] = [%r, %r, %r]
import pyarrow as pa
from polars_strsim import _native
from oracle import oracle
from bench_support import plugin_driver, workloads

L = _native.lib()
assert L.strsim_b200_device_count() >= 2
MEASURES = list(oracle.MEASURES)

def check(A, B, a, b, what):
    outs, valid, nulls, ints = _native.compute_host_multi(MEASURES, A, B, debug=True)
    for m, v, i in zip(MEASURES, outs, ints):
        ref, rv, ri = oracle.batch(m, a, b)
        assert (valid == rv).all() and nulls == int((~rv).sum()), (what, m, "null mask")
        assert (v[rv].view(np.uint64) == ref[rv].view(np.uint64)).all(), (what, m)
        assert (i[rv] == ri[rv]).all(), (what, m, "ints")

# 1. C3-like rows (mixed scripts, 5 %% nulls), different chunkings on the two sides, pageable memory
n = 400_000
A, B = workloads.make_pairs(3, n, uneven_b=True)
a, b = A.to_pylist(), B.to_pylist()
check(A, B, a, b, "C3 uneven chunks")
# 2. literal operands in both orientations
lit = pa.array(["josé maría"], type=pa.string_view())
check(A, lit, a, ["josé maría"] * n, "literal right")
check(lit, B, ["josé maría"] * n, b, "literal left")
try:
    _native.compute_host("jaro", A, pa.array([None], type=pa.string_view()))
    raise AssertionError("a null literal must fail the call")
except _native.StrsimError as exc:
    assert "literal operand is null" in str(exc)
# 3. a column that looks sequential but holds one view pointing far back (row k repeats row j): the shard
#    that owns row k uploaded only its own stretch of the data, finds the view outside it and the call
#    repeats that shard with whole buffers -- results must not change
A2, B2 = workloads.make_pairs(2, n)
views = np.frombuffer(A2.buffers()[1], dtype=np.int32).reshape(-1, 4).copy()
long_rows = np.nonzero(views[:, 0] > 12)[0]
j, k = int(long_rows[10]), int(long_rows[long_rows > (3 * n) // 4][7])
views[k] = views[j]
A3 = pa.Array.from_buffers(pa.string_view(), n, [None, pa.py_buffer(views.tobytes()), A2.buffers()[2]])
a3 = A3.to_pylist()
assert a3[k] == a3[j]
check(A3, B2, a3, B2.to_pylist(), "view outside the shard's stretch")
# 4. the plugin path: five calls, one (sharded) upload, composite columns kept in the cache
L.strsim_b200_cache_stats.argtypes = [ctypes.POINTER(ctypes.c_int64)]
L.strsim_b200_cache_stats.restype = None
def stats():
    out = (ctypes.c_int64 * 4)()
    L.strsim_b200_cache_stats(out)
    return list(out)
plugin_driver.cache_clear()
h0, m0, _, _ = stats()
for m in MEASURES:
    r = plugin_driver.call(m, A, B)
    got = np.concatenate(r.values())
    valid = r.validity()
    ref, rv, _ = oracle.batch(m, a, b)
    assert (valid == rv).all() and r.null_count == int((~rv).sum())
    assert (got[rv].view(np.uint64) == ref[rv].view(np.uint64)).all(), m
    r.release()
h, mi, cols, nbytes = stats()
assert (h - h0, mi - m0, cols) == (8, 2, 2), (h - h0, mi - m0, cols)
plugin_driver.cache_clear()
# 5. a small call stays on one device
check(A.slice(0, 1000), B.slice(0, 1000), a[:1000], b[:1000], "small call")
print("ok")
"""


@pytest.mark.gpu
def test_one_call_sharded_over_two_devices():
    sys.path[:0] = [str(ROOT), str(ROOT / "polars-strsim_b200")]
    from polars_strsim import _native

    if _native.lib().strsim_b200_device_count() < 2:
        pytest.skip("needs two CUDA devices (gpurun --gpus 2)")
    code = CODE % (str(ROOT), str(ROOT / "polars-strsim_b200"), str(ROOT / "tests"))
    env = dict(os.environ, STRSIM_B200_DEVICES="0,1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "ok" in out.stdout, (out.stdout[-2000:], out.stderr[-4000:])
