"""Synthetic workloads of BASELINE.json as pyarrow string_view arrays.

Bench / test support: drives bench_support/gen.c (seeded, thread-count independent) and wraps the
buffers it fills -- optionally pinned host memory -- as Arrow arrays without copying.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np
import pyarrow as pa

HERE = Path(__file__).resolve().parent
LIB = HERE / "libgen.so"

SEEDS = {2: 0xC2, 3: 0xC3, 4: 0xC4, 5: 0xC5, 6: 0xC6, 7: 0xC7, 8: 0xC8, 9: 0xC9}
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not LIB.exists() or LIB.stat().st_mtime < (HERE / "gen.c").stat().st_mtime:
            subprocess.run(["make", "-C", str(HERE), "-B"], check=True, capture_output=True)
        L = ctypes.CDLL(str(LIB))
        L.gen_pairs.restype = ctypes.c_int
        L.gen_pairs.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64, ctypes.c_double,
                                ctypes.c_int] + [ctypes.c_void_p] * 6 + [ctypes.POINTER(ctypes.c_int64)]
        _lib = L
    return _lib


def _alloc(nbytes: int, pinned: bool):
    """uint8 buffer of nbytes (+64 spare) as (owner, address)."""
    if pinned:
        import torch

        t = torch.empty(nbytes + 64, dtype=torch.uint8, pin_memory=True)
        return t, t.data_ptr()
    a = np.empty(nbytes + 64, dtype=np.uint8)
    return a, a.ctypes.data


def make_chunk(config: int, n: int, row_base: int = 0, null_p: float = 0.0, pinned: bool = False,
               seed: int | None = None, threads: int | None = None):
    """One chunk of `n` pairs -> (A, B) pyarrow string_view arrays (+ host bytes held)."""
    L = lib()
    seed = SEEDS.get(config, config) if seed is None else seed
    if not threads:  # the ranks of a one-process-per-GPU launch share the host's cores
        threads = max(1, len(os.sched_getaffinity(0)) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))
    gen_config = 2 if config == 5 else config
    sizes = (ctypes.c_int64 * 2)()
    va, pva = _alloc(16 * n, pinned)
    vb, pvb = _alloc(16 * n, pinned)
    rc = L.gen_pairs(gen_config, seed, row_base, n, null_p, threads, None, None, None, None, None, None, sizes)
    if rc != 0:
        raise ValueError("more than 2 GiB of out-of-line bytes in one chunk: use more chunks")
    da, pda = _alloc(sizes[0], pinned)
    db, pdb = _alloc(sizes[1], pinned)
    if null_p > 0:
        na, pna = _alloc((n + 7) // 8, pinned)
        nb, pnb = _alloc((n + 7) // 8, pinned)
    else:
        na = nb = None
        pna = pnb = None
    rc = L.gen_pairs(gen_config, seed, row_base, n, null_p, threads, pva, pda, pna, pvb, pdb, pnb, sizes)
    assert rc == 0

    def arr(views, pviews, data, pdata, size, valid, pvalid):
        bufs = [pa.foreign_buffer(pvalid, (n + 7) // 8, base=valid) if valid is not None else None,
                pa.foreign_buffer(pviews, 16 * n, base=views),
                pa.foreign_buffer(pdata, size, base=data)]
        return pa.Array.from_buffers(pa.string_view(), n, bufs)

    A = arr(va, pva, da, pda, sizes[0], na, pna)
    B = arr(vb, pvb, db, pdb, sizes[1], nb, pnb)
    return A, B


def make_pairs(config: int, n: int, row_base: int = 0, null_p: float | None = None, pinned: bool = False,
               n_chunks: int | None = None, uneven_b: bool = False, seed: int | None = None):
    """(A, B) for a BASELINE config: C2 10M ASCII, C3 mixed Unicode + 5 % nulls, C4 long text, C5 = C2
    generator with seed 0xC5.  C4 is split so that every data buffer stays below 2 GiB."""
    if null_p is None:
        null_p = 0.05 if config == 3 else 0.0
    if n_chunks is None:
        n_chunks = max(1, -(-n // 400_000)) if config == 4 else 1
    bounds = [n * i // n_chunks for i in range(n_chunks + 1)]
    As, Bs = [], []
    for lo, hi in zip(bounds, bounds[1:]):
        a, b = make_chunk(config, hi - lo, row_base + lo, null_p, pinned, seed)
        As.append(a)
        Bs.append(b)
    if n_chunks == 1 and not uneven_b:
        return As[0], Bs[0]
    A = pa.chunked_array(As)
    B = pa.chunked_array(Bs)
    if uneven_b:  # different chunking on the two sides (C3: "several chunks of unequal size per column")
        cuts = sorted({0, n // 7, n // 3, (2 * n) // 3 + 1, n})
        flat = pa.concat_arrays(Bs) if len(Bs) > 1 else Bs[0]
        B = pa.chunked_array([flat.slice(lo, hi - lo) for lo, hi in zip(cuts, cuts[1:])])
    return A, B


def algorithmic_bytes(A, B, measures: int = 1) -> int:
    """per pair 2x16 B of views + out-of-line payload (byte length > 12) + validity
    bits in (if a bitmap exists) + 8 B out (+ 1 bit out validity if any input bitmap)."""
    total = 0
    any_valid = False
    n = len(A)
    for col in (A, B):
        for ch in (col.chunks if isinstance(col, pa.ChunkedArray) else [col]):
            bufs = ch.buffers()
            v = np.frombuffer(bufs[1], dtype=np.int32)[4 * ch.offset: 4 * (ch.offset + len(ch))].reshape(-1, 4)
            lens = v[:, 0].astype(np.int64)
            total += 16 * len(ch) + int(lens[lens > 12].sum())
            if bufs[0] is not None:
                total += (len(ch) + 7) // 8
                any_valid = True
    total += 8 * n + ((n + 7) // 8 if any_valid else 0)
    return total * measures
