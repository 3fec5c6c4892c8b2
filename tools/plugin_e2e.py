#!/usr/bin/env python
"""The README scenario through the real plugin symbols: five `_polars_plugin_<measure>` calls over the
same two columns (fabricated polars-ffi SeriesExports, as in tests/test_abi.py), timed end to end with
host buffers -- with and without the library's device-side cache of plugin inputs.

    python tools/plugin_e2e.py [rows]            # run on a GPU box; STRSIM_B200_CACHE=0 disables the cache
"""
import ctypes
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "polars-strsim_b200"), str(ROOT / "tests")]

from bench_support import workloads  # noqa: E402
from polars_strsim import _native  # noqa: E402
from test_abi import SeriesExport, make_series  # noqa: E402

MEASURES = ("levenshtein", "jaro", "jaro_winkler", "jaccard", "sorensen_dice")


def main():
    args = [x for x in sys.argv[1:] if not x.startswith("--")]
    pageable = "--pageable" in sys.argv  # inputs in ordinary (pageable) memory, as Polars hands them over
    n = int(args[0]) if args else 10_000_000
    L = _native.lib()
    A, B = workloads.make_pairs(2, n, pinned=not pageable)
    L.strsim_b200_cache_clear.restype = None

    def five_calls():
        sums = []
        for m in MEASURES:
            released, keep = [], []
            inputs = (SeriesExport * 2)()
            inputs[0], _ = make_series(A, released, keep)
            inputs[1], _ = make_series(B, released, keep)
            ret = SeriesExport()
            getattr(L, f"_polars_plugin_{m}")(inputs, ctypes.c_size_t(2), None, ctypes.c_size_t(0), ctypes.byref(ret), None)
            assert ret.private_data, L.strsim_b200_last_error()
            arr = _native.ArrowArray.from_address(ret.arrays[0])
            vals = (ctypes.c_double * 4).from_address(arr.buffers[1])
            sums.append(vals[0] + vals[3])
            arr.release(ctypes.byref(arr))
            ret.release(ctypes.byref(ret))
        return sums

    five_calls()
    L.strsim_b200_cache_clear()
    times = []
    for _ in range(3):
        L.strsim_b200_cache_clear()  # every repetition starts cold, like a fresh query
        t0 = time.perf_counter()
        five_calls()
        times.append((time.perf_counter() - t0) * 1e3)
    print(json.dumps({"rows": n, "cache": os.environ.get("STRSIM_B200_CACHE", "1"),
                      "inputs": "pageable" if pageable else "pinned",
                      "five_plugin_calls_ms": min(times), "all_ms": times,
                      "note": "results land in pageable memory malloc'ed by the plugin (Arrow result buffers)"}))


if __name__ == "__main__":
    main()
