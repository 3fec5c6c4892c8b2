"""Row-range sharding: host-side logic on CPU, world size 2 over gloo.

The per-rank compute is stood in for by the oracle (this is test code; the product path needs a
GPU) -- what is under test is the partition, the per-rank slicing of Arrow columns (non-zero
offsets) and the rank-ordered concatenation, i.e. everything bench.py / a multi-GPU caller does
around the single-device entry point."""
import os
import random
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_split_offsets_matches_reference_rule():
    from polars_strsim.sharding import shard_for_rank, split_offsets

    assert split_offsets(10, 1) == [(0, 10)]
    assert split_offsets(10, 4) == [(0, 2), (2, 2), (4, 2), (6, 4)]  # last takes the remainder
    assert split_offsets(3, 8)[-1] == (0, 3) and split_offsets(3, 8)[0] == (0, 0)
    for n in (0, 1, 7, 64, 1000):
        for w in (1, 2, 3, 8):
            parts = split_offsets(n, w)
            assert sum(l for _, l in parts) == n
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert shard_for_rank(n, w, w - 1) == parts[-1]


def test_in_call_shard_cuts_follow_split_offsets_on_64_row_units():
    """The partition one sharded host call uses (STRSIM_B200_DEVICES; include/strsim_b200.h:
    strsim_b200_shard_cuts): the reference's split_offsets (strsim.rs:21-39) applied to units of 64 rows,
    so that no validity byte is shared between two devices; the last range takes the remainder."""
    import ctypes

    sys.path[:0] = [str(ROOT), str(ROOT / "polars-strsim_b200")]
    from polars_strsim import _native
    from polars_strsim.sharding import split_offsets

    L = _native.lib()
    L.strsim_b200_shard_cuts.restype = None
    L.strsim_b200_shard_cuts.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.POINTER(ctypes.c_int64)]
    for n in (0, 1, 63, 64, 65, 65536, 1_000_003, 10_000_000, 2**31 + 5):
        for g in (1, 2, 3, 4, 8):
            cuts = (ctypes.c_int64 * (g + 1))()
            L.strsim_b200_shard_cuts(n, g, cuts)
            cuts = list(cuts)
            assert cuts[0] == 0 and cuts[-1] == n and all(x <= y for x, y in zip(cuts, cuts[1:]))
            assert all(c % 64 == 0 for c in cuts[:-1])
            units = split_offsets((n + 63) // 64, g)
            assert cuts[:-1] == [off * 64 for off, _ in units]
            sizes = [y - x for x, y in zip(cuts, cuts[1:])]
            assert len(set(sizes[:-1])) <= 1 and (g == 1 or sizes[-1] >= sizes[0] - 63)


def _worker(rank, world, port, out_dir):
    sys.path[:0] = [str(ROOT), str(ROOT / "polars-strsim_b200"), str(ROOT / "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import pyarrow as pa
    import torch
    import torch.distributed as dist

    from oracle import oracle
    from polars_strsim.sharding import compute_sharded
    from test_oracle import rand_pair

    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = random.Random(42)  # every rank builds the same logical frame
    pairs = [rand_pair(rng, 30) for _ in range(3001)]
    a = [None if rng.random() < 0.05 else p[0] for p in pairs]
    b = [None if rng.random() < 0.05 else p[1] for p in pairs]
    A = pa.chunked_array([pa.array(a[:1000], type=pa.string_view()), pa.array(a[1000:], type=pa.string_view())])
    B = pa.array(b, type=pa.string_view())

    def compute(measure, x, y):
        x = x.combine_chunks() if isinstance(x, pa.ChunkedArray) else x
        return oracle.batch_views(measure, x, y, n_threads=2)

    ok = True
    for measure in ("levenshtein", "jaro_winkler", "sorensen_dice"):
        off, vals, valid = compute_sharded(measure, A, B, world, rank, compute)
        # host-side concatenation in rank order (gather to rank 0), no reduction of any kind
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([len(vals)], dtype=torch.int64))
        mx = int(max(s.item() for s in sizes))
        pad = torch.zeros(mx, dtype=torch.float64)
        pad[: len(vals)] = torch.from_numpy(np.where(valid, vals, -1.0))
        gathered = [torch.zeros(mx, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, pad)
        full = np.concatenate([g.numpy()[: int(s.item())] for g, s in zip(gathered, sizes)])
        ref, ref_valid, _ = oracle.batch(measure, a, b)
        ok = ok and len(full) == len(ref) and bool((full == np.where(ref_valid, ref, -1.0)).all())
    # a scalar literal is broadcast on every rank
    off, vals, valid = compute_sharded("jaro", A, "smith", world, rank,
                                       lambda m, x, y: compute(m, x, pa.array([y], type=pa.string_view())))
    ref, ref_valid, _ = oracle.batch("jaro", a[off: off + len(vals)], ["smith"] * len(vals))
    ok = ok and bool((vals[valid] == ref[ref_valid]).all())
    Path(out_dir, f"rank{rank}.ok").write_text("1" if ok else "0")
    dist.destroy_process_group()


def test_two_rank_row_sharding_gloo(tmp_path):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert [(tmp_path / f"rank{r}.ok").read_text() for r in range(2)] == ["1", "1"]
