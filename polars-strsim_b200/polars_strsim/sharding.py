"""Row-range sharding: the only multi-GPU strategy of this path (rows are independent).

`split_offsets` is the reference's static partition (/root/reference/src/expressions/strsim.rs:21-39):
n contiguous ranges of len // n rows, the last one takes the remainder.  Each GPU (one process per
GPU, or one host thread per GPU inside the plugin) owns one range; results are concatenated on the
host in rank order -- no collective, no NCCL on the data path.
"""
from __future__ import annotations


def split_offsets(length: int, n: int):
    """[(offset, len)] * n exactly as strsim.rs:21-39."""
    if n == 1:
        return [(0, length)]
    chunk = length // n
    out = []
    for p in range(n):
        off = p * chunk
        out.append((off, length - off if p == n - 1 else chunk))
    return out


def shard_for_rank(length: int, world_size: int, rank: int):
    """(offset, len) of `rank`'s row range."""
    return split_offsets(length, world_size)[rank]


def slice_column(col, offset: int, length: int):
    """Zero-copy row range of a pyarrow Array / ChunkedArray (a scalar literal is returned as is)."""
    if isinstance(col, (str, bytes)) or col is None or len(col) == 1:
        return col
    return col.slice(offset, length)


def compute_sharded(measure, a, b, world_size: int, rank: int, compute):
    """This rank's slice of measure(a, b): returns (offset, values, valid).  `compute(measure, a, b)`
    is the single-device entry point (polars_strsim._native.compute_host on a GPU)."""
    n = max(len(a) if hasattr(a, "__len__") and not isinstance(a, (str, bytes)) else 1,
            len(b) if hasattr(b, "__len__") and not isinstance(b, (str, bytes)) else 1)
    off, ln = shard_for_rank(n, world_size, rank)
    vals, valid, *_ = compute(measure, slice_column(a, off, ln), slice_column(b, off, ln))
    return off, vals, valid
