"""Debug aid: fused vs single-measure results on rows of 33..64 ASCII bytes (the 64-bit plane kernel)."""
import sys
from pathlib import Path

import numpy as np
import pyarrow as pa

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "polars-strsim_b200"))
from oracle import oracle  # noqa: E402
from polars_strsim import _native  # noqa: E402
from bench_support import workloads  # noqa: E402

A, B = workloads.make_pairs(7, 5000)
a, b = A.to_pylist(), B.to_pylist()
names = list(oracle.MEASURES)
outs, valid, nulls, ints = _native.compute_host_multi(names, A, B, debug=True)
for m, got, gi in zip(names, outs, ints):
    ref, rv, ri = oracle.batch(m, a, b)
    bad = np.nonzero(got.view(np.uint64) != ref.view(np.uint64))[0]
    print(m, "fused mismatches", bad.size, "of", len(a))
    for i in bad[:4]:
        print("   ", i, len(a[i]), len(b[i]), repr(a[i]), repr(b[i]), got[i], ref[i], gi[i], ri[i])
    v1, _, _, i1 = _native.compute_host(m, A, B, debug=True)
    bad1 = np.nonzero(v1.view(np.uint64) != ref.view(np.uint64))[0]
    print(m, "single mismatches", bad1.size)
