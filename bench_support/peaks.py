"""Measured integer issue-rate ceilings (bench_support/peaks.cu) in warp instructions / clock / SM."""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB = HERE / "libpeaks.so"
MODES = {"lop3": 0, "imad": 2, "lop3_imad_mix": 3}


def measure(sm_mhz: float, iters: int = 4096, reps: int = 5) -> dict:
    """sm_mhz: SM clock under load (bench.py samples it with nvidia-smi).  Returns per mode the issue
    rate in warp instructions per clock per SM, and the same in warp instructions per second."""
    if not LIB.exists():
        subprocess.run(["make", "-C", str(HERE)], check=True, capture_output=True)
    L = ctypes.CDLL(str(LIB))
    L.int_peak_run.restype = ctypes.c_int
    L.int_peak_run.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                               ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)]
    out = {}
    for name, mode in MODES.items():
        ms, wi, sms = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
        rc = L.int_peak_run(mode, iters, reps, ctypes.byref(ms), ctypes.byref(wi), ctypes.byref(sms))
        if rc != 0:
            raise RuntimeError(f"int_peak_run({name}) failed: {rc}")
        per_s = wi.value * reps / (ms.value * 1e-3)
        out[name] = {"warp_inst_per_s": per_s, "per_clk_per_sm": per_s / (sms.value * sm_mhz * 1e6)}
    out["sm_mhz"] = sm_mhz
    out["sm_count"] = sms.value
    return out
