#!/usr/bin/env python
"""Runs ONE measure (or the fused set) over a workload's resident columns -- the target of an ncu capture.

    ncu --set full --import-source on -k regex:long_pair -c 1 -o gpurun_out/x python tools/prof_one.py T1 jaro 2000000
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "polars-strsim_b200")]
import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from bench_support import workloads  # noqa: E402
from polars_strsim import _native  # noqa: E402

wl = WORKLOADS[sys.argv[1]]
what = sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else wl["rows"]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
A, B = workloads.make_pairs(wl["config"], n, uneven_b=(wl["config"] == 3))
colA, colB = _native.DeviceColumn(A), _native.DeviceColumn(B)
measures = list(wl["measures"]) if what == "fused" else [what]
outs = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in measures]
val = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
st = torch.cuda.Stream()
for _ in range(reps):
    if len(measures) > 1:
        _native.compute_device_multi(measures, colA, colB, [o.data_ptr() for o in outs], val.data_ptr(), None, st.cuda_stream)
    else:
        _native.compute_device(measures[0], colA, colB, outs[0].data_ptr(), val.data_ptr(), 0, st.cuda_stream)
torch.cuda.synchronize()
print("ok", [float(o.sum().item()) for o in outs])
