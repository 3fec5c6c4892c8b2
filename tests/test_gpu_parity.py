"""GPU parity: the CUDA path (through the C ABI) against the oracle.  Run with `-m gpu` on a B200.

Bar (BASELINE.json north_star): integer intermediates bit-exact, f64 results bit-identical
(target 0 ulp; the contract allows 1 ulp), identical null masks.
"""
import json
import os
import random
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from test_oracle import GOLDEN, load_fixture, rand_pair

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def native():
    from polars_strsim import _native

    assert _native.lib().strsim_b200_device_count() > 0, "no CUDA device: these tests need the B200"
    return _native


def sv(values):
    import pyarrow as pa

    return pa.array(values, type=pa.string_view())


def ulp_diff(x, y):
    xi = x.view(np.int64)
    yi = y.view(np.int64)
    return np.abs(xi - yi)


def check(native, oracle, measure, a, b, A=None, B=None, ref_a=None, ref_b=None):
    """a, b: python lists (the logical rows); A, B: arrow inputs if they differ from sv(a), sv(b)."""
    A = sv(a) if A is None else A
    B = sv(b) if B is None else B
    ra = a if ref_a is None else ref_a
    rb = b if ref_b is None else ref_b
    ref, ref_valid, ref_ints = oracle.batch(measure, ra, rb)
    vals, valid, nulls, ints = native.compute_host(measure, A, B, debug=True)
    assert len(vals) == len(ref)
    assert (valid == ref_valid).all(), "null masks differ"
    assert nulls == int((~ref_valid).sum())
    bad = np.nonzero(valid & (vals.view(np.uint64) != ref.view(np.uint64)))[0]
    assert bad.size == 0, (measure, ra[bad[0]], rb[bad[0]], vals[bad[0]], ref[bad[0]], ints[bad[0]], ref_ints[bad[0]])
    ibad = np.nonzero(valid & (ints != ref_ints).any(axis=1))[0]
    assert ibad.size == 0, (measure, ra[ibad[0]], rb[ibad[0]], ints[ibad[0]], ref_ints[ibad[0]])
    return vals, valid


def test_reference_golden_vectors(native, oracle):
    """All 1115 known-answer vectors of the reference's own tests (strsim.rs:371-1534)."""
    fx = load_fixture()
    for measure in oracle.MEASURES:
        rows = [r for r in fx if r[0] == measure]
        vals, valid = check(native, oracle, measure, [r[1] for r in rows], [r[2] for r in rows])
        exp = np.array([r[3] for r in rows])
        exact = np.array([r[4] for r in rows])
        assert valid.all()
        assert (np.abs(vals - exp) < 1e-8).all()          # the reference's own tolerance, strsim.rs:349
        assert (vals.view(np.uint64) == exact.view(np.uint64)).all()  # and bit-exact vs the fixture


def test_readme_table(native, oracle):
    table = json.loads((GOLDEN / "readme_table.json").read_text())
    for measure in oracle.MEASURES:
        vals, valid, nulls = native.compute_host(measure, sv(table["name_a"]), sv(table["name_b"]))
        assert nulls == 2
        for v, ok, printed, hx in zip(vals, valid, table["printed"][measure], table["oracle_hex"][measure]):
            assert ok == (printed is not None)
            if ok:
                assert abs(v - printed) < 5e-7 and v == float.fromhex(hx)


def test_random_short_mixed_scripts(native, oracle):
    rng = random.Random(2024)
    pairs = [rand_pair(rng, 24) for _ in range(30000)]
    a, b = [p[0] for p in pairs], [p[1] for p in pairs]
    for measure in oracle.MEASURES:
        check(native, oracle, measure, a, b)


def test_length_boundaries(native, oracle):
    """12/13 bytes (inline vs out-of-line view), 32/33 and 64/65 (mask word boundaries), 1-4 byte UTF-8."""
    rng = random.Random(77)
    a, b = [], []
    lens = [0, 1, 2, 3, 4, 11, 12, 13, 14, 15, 16, 17, 31, 32, 33, 34, 63, 64, 65, 66, 100]
    for la in lens:
        for lb in lens:
            for alpha in ("ab", "abcdefghijklmnopqrstuvwxyz"):
                x = "".join(rng.choice(alpha) for _ in range(la))
                y = "".join(rng.choice(alpha) for _ in range(lb))
                a += [x, x]
                b += [y, x[: lb] if lb <= la else x + y[la:]]
    for ch in ("é", "ß", "日", "\U0001f600"):
        w = len(ch.encode())
        for la in (1, 2, 3, 12 // w, 12 // w + 1, 32 // w, 32 // w + 1, 64 // w, 64 // w + 1):
            for lb in (1, 2, 12 // w, 32 // w, 64 // w + 1):
                a += [ch * la, ch * la, "x" + ch * la]
                b += [ch * lb, ("z" * lb), ch * lb + "y"]
    for measure in oracle.MEASURES:
        check(native, oracle, measure, a, b)


def test_medium_and_long_strings(native, oracle):
    rng = random.Random(31337)
    pairs = [rand_pair(rng, 64) for _ in range(4000)] + [rand_pair(rng, 200) for _ in range(600)]
    pairs += [rand_pair(rng, 900) for _ in range(40)]
    a, b = [p[0] for p in pairs], [p[1] for p in pairs]
    for measure in oracle.MEASURES:
        check(native, oracle, measure, a, b)
    assert sum(native.last_overflow()) > 0  # the overflow kernels really ran


def test_wide_rows_longer_than_the_column_mean(native, oracle):
    """A general column of short Latin-1 rows with one row in ten a long CJK string: the launch over the wide
    pairs (short_kernel.cuh, second launch of launch_general) sizes its staging area from the COLUMN's mean
    payload, so its tiles -- wide pairs only -- overflow it.  Those rows must finish in the long kernel, not
    vanish (they did: the second launch assumed the first had listed them)."""
    rng = random.Random(4242)
    cjk = [chr(c) for c in range(0x4E00, 0x4E40)]
    a, b = [], []
    for r in range(60000):
        if r % 10 == 3:
            x = "".join(rng.choice(cjk) for _ in range(rng.randint(7, 10)))
            y = list(x)
            y[rng.randrange(len(y))] = rng.choice(cjk)
            y = "".join(y[: rng.randint(6, len(y))])
        else:
            x = "".join(rng.choice("abcdé") for _ in range(rng.randint(1, 5)))
            y = "".join(rng.choice("abcdé") for _ in range(rng.randint(1, 5)))
        a.append(x)
        b.append(y)
    for measure in ("levenshtein", "jaro_winkler", "sorensen_dice"):
        check(native, oracle, measure, a, b)
    check_multi(native, oracle, list(range(5)), a, b)


def test_wide_rows_of_ascii_columns(native, oracle):
    """Rows of 65..320 bytes of ASCII-only columns: one pair per thread with masks of ten words
    (wide_mask.cuh, host.cu: finish_wide) -- between the 64-bit plane launch and the warp-per-pair kernels,
    which still take what is longer.  Lengths around every 32-character word boundary, near-duplicates,
    moved blocks, tiny alphabets, equal pairs, one side short or empty, nulls; every measure alone, all
    five fused, and group subsets."""
    rng = random.Random(777)
    alpha = "abcdefghijklmnopqrstuvwxyz ABCDEFGH0123456789-,.'#/"
    a, b = [], []
    for _ in range(6000):
        n = rng.choice([rng.randint(65, 320), rng.randint(65, 320), rng.randint(1, 64), rng.randint(321, 420)])
        al = alpha if rng.random() < 0.8 else "ab"
        x = "".join(rng.choice(al) for _ in range(n))
        r = rng.random()
        if r < 0.65:
            y = list(x)
            for _ in range(rng.randint(0, 10)):
                k = rng.randrange(len(y))
                op = rng.random()
                if op < 0.35:
                    y[k] = rng.choice(al)
                elif op < 0.6:
                    y.insert(k, rng.choice(al))
                elif op < 0.85 and len(y) > 1:
                    del y[k]
                elif len(y) > 40:
                    blk = y[k:k + 12]
                    del y[k:k + 12]
                    q = rng.randrange(len(y) + 1)
                    y[q:q] = blk
            y = "".join(y)
        elif r < 0.85:
            y = "".join(rng.choice(al) for _ in range(rng.randint(0, 340)))
        elif r < 0.93:
            y = x
        else:
            y = x[: rng.randint(0, 5)]
        if rng.random() < 0.03:
            x = None
        if rng.random() < 0.03:
            y = None
        a.append(x)
        b.append(y)
    for la in (64, 65, 96, 97, 128, 129, 160, 161, 288, 289, 319, 320, 321):
        for lb in (0, 1, 33, 65, 160, 319, 320, 321):
            a += ["a" * la, "".join(rng.choice(alpha) for _ in range(la))]
            b += ["b" * lb, "".join(rng.choice(alpha) for _ in range(lb))]
    for measure in oracle.MEASURES:
        check(native, oracle, measure, a, b)
    assert native.last_overflow()[1] > 0
    check_multi(native, oracle, list(range(5)), a, b)
    for ids in ([1, 2], [3, 4], [0, 3], [0, 1]):
        check_multi(native, oracle, ids, a, b)


def test_long_rows_one_pair_per_warp(native, oracle):
    """Rows above 64 bytes of Jaro / Jaro-Winkler / Jaccard / Sorensen-Dice (long_pair_kernel.cuh: one pair per
    warp -- windowed matching by ballot over 32 positions of b per step, transpositions on compacted flag
    sets, multiset intersection through a per-pair hash table; the reference treats every length alike,
    strsim.rs:208-219,297-305): tiny alphabets (every window full of candidates, long runs of equal
    characters), anagrams, reversals, near-duplicates, unrelated pairs, 1-4 byte UTF-8, lengths around the
    32-position word boundaries up to 2500 characters, one side short or empty, nulls -- single-measure
    launches, the fused launch, and the one-thread-per-pair kernel as an independent second opinion."""
    rng = random.Random(4242)
    a, b = [], []
    alphabets = ["ab", "abc", "abcdefghijklmnopqrstuvwxyz ", "aéß日\U0001f600xyz", "日本語中文字漢"]
    for n in (65, 66, 95, 96, 97, 127, 128, 129, 200, 300, 511, 1000, 2500):
        for alpha in alphabets:
            x = "".join(rng.choice(alpha) for _ in range(n))
            y = list(x)
            for _ in range(max(1, n // 12)):  # a few edits
                k = rng.randrange(len(y))
                op = rng.random()
                if op < 0.4:
                    y[k] = rng.choice(alpha)
                elif op < 0.7:
                    y.insert(k, rng.choice(alpha))
                elif len(y) > 1:
                    del y[k]
            shuffled = list(x)
            rng.shuffle(shuffled)
            z = "".join(rng.choice(alpha) for _ in range(rng.randint(1, n + 40)))
            a += [x, x, x, x, x, x, x[: n // 2], "", x, None, x + "q"]
            b += ["".join(y), "".join(shuffled), x[::-1], z, x, x[:3], x, x, "", x, None]
    a += ["x" * 300, "ab" * 150, "a" * 70 + "b" * 70]
    b += ["x" * 150 + "y" * 150, "ba" * 150, "b" * 70 + "a" * 70]
    for measure in oracle.MEASURES:
        check(native, oracle, measure, a, b)
    assert native.last_overflow()[1] > 0  # the long-row kernels really ran
    check_multi(native, oracle, list(range(5)), a, b)
    check_multi(native, oracle, [2, 4], a, b)


def test_nulls_slices_chunks_broadcast(native, oracle):
    import pyarrow as pa

    rng = random.Random(5)
    n = 7001
    a, b = [], []
    for _ in range(n):
        x, y = rand_pair(rng, 30)
        a.append(None if rng.random() < 0.05 else x)
        b.append(None if rng.random() < 0.05 else y)
    for measure in oracle.MEASURES:
        check(native, oracle, measure, a, b)
    # sliced arrays: non-zero ArrowArray.offset on views AND validity
    A, B = sv(a), sv(b)
    check(native, oracle, "jaro_winkler", a[13:5013], b[101:5101], A=A.slice(13, 5000), B=B.slice(101, 5000))
    # different chunkings on the two sides
    cuts_a, cuts_b = [0, 1, 64, 1000, 4097, n], [0, 333, 2048, 2049, n]
    CA = pa.chunked_array([A.slice(lo, hi - lo) for lo, hi in zip(cuts_a, cuts_a[1:])])
    CB = pa.chunked_array([B.slice(lo, hi - lo) for lo, hi in zip(cuts_b, cuts_b[1:])])
    for measure in ("levenshtein", "sorensen_dice"):
        check(native, oracle, measure, a, b, A=CA, B=CB)
    # scalar broadcast, both orientations (the left-literal case returns n rows, not the
    # reference's 1-row quirk of strsim.rs:73)
    for measure in oracle.MEASURES:
        check(native, oracle, measure, a, ["smith"] * n, B=sv(["smith"]))
        check(native, oracle, measure, ["josé maría"] * n, b, A=sv(["josé maría"]))
    # a null literal fails the query (the reference panics on it, strsim.rs:62,65); a one-row frame is not a
    # literal: its nulls propagate like any other row's
    for A_, B_ in ((sv(a), sv([None])), (sv([None]), sv(b))):
        with pytest.raises(native.StrsimError, match="literal operand is null"):
            native.compute_host("jaro", A_, B_)
    check(native, oracle, "jaro", ["x"], [None])
    check(native, oracle, "jaro", [None], ["x"])
    # all null, single row, zero rows
    check(native, oracle, "jaccard", [None] * 100, b[:100])
    check(native, oracle, "levenshtein", ["x"], ["y"])
    vals, valid, nulls = native.compute_host("jaro", sv([]), sv([]))
    assert len(vals) == 0 and nulls == 0
    with pytest.raises(native.StrsimError, match="same length"):
        native.compute_host("jaro", sv(a[:10]), sv(b[:11]))


def test_scattered_views_take_the_gather_path(native, oracle):
    """Views that do not reference one contiguous span (a take() result) and several data buffers."""
    import pyarrow as pa

    rng = random.Random(11)
    base = ["".join(rng.choice("abcdefghij") for _ in range(rng.randint(13, 30))) for _ in range(20000)]
    idx = [rng.randrange(len(base)) for _ in range(6000)]
    src = sv(base)
    bufs = src.buffers()
    views = np.frombuffer(bufs[1], dtype=np.int32).reshape(-1, 4)
    A = pa.Array.from_buffers(pa.string_view(), len(idx),
                              [None, pa.py_buffer(np.ascontiguousarray(views[idx]).tobytes())] + bufs[2:])
    assert A.to_pylist() == [base[i] for i in idx]
    a = [base[i] for i in idx]
    b = ["".join(rng.choice("abcdefghij") for _ in range(rng.randint(0, 30))) for _ in idx]
    for measure in oracle.MEASURES:
        check(native, oracle, measure, a, b, A=A)
    big = pa.concat_arrays([sv(base[:10000]), sv(base[10000:])])  # >= 2 variadic buffers
    check(native, oracle, "levenshtein", base, list(reversed(base)), A=big)


def test_arrow_c_data_interface_and_offset_layouts(native, oracle):
    import pyarrow as pa
    from polars_strsim import arrow

    rng = random.Random(3)
    pairs = [rand_pair(rng, 40) for _ in range(3000)]
    a = [None if rng.random() < 0.1 else p[0] for p in pairs]
    b = [p[1] for p in pairs]
    for measure in oracle.MEASURES:
        ref, ref_valid, _ = oracle.batch(measure, a, b)
        for typ in (pa.string_view(), pa.string(), pa.large_string()):
            out = getattr(arrow, measure)(pa.array(a, type=typ), pa.array(b, type=typ))
            assert out.type == pa.float64() and len(out) == len(a)
            got_valid = np.array([v is not None for v in out.to_pylist()])
            assert (got_valid == ref_valid).all()
            got = np.array([0.0 if v is None else v for v in out.to_pylist()])
            assert (got[ref_valid] == ref[ref_valid]).all()
    out = arrow.jaro_winkler(pa.chunked_array([pa.array(a[:100]), pa.array(a[100:])]).slice(7), pa.array(b[7:]))
    ref, ref_valid, _ = oracle.batch("jaro_winkler", a[7:], b[7:])
    got = np.array([0.0 if v is None else v for v in out.to_pylist()])
    assert (got[ref_valid] == ref[ref_valid]).all()
    with pytest.raises(native.StrsimError, match="expected `String`"):
        arrow.jaro(pa.array([1, 2]), pa.array(["a", "b"]))


def test_fallback_kernel_cross_check(oracle):
    """STRSIM_B200_FORCE_GENERIC=1 sends every row through the independent textbook kernel."""
    code = r"""
import sys, random, numpy as np
sys.path[:0] = [%r, %r, %r]
import pyarrow as pa
from polars_strsim import _native
from oracle import oracle
from test_oracle import rand_pair
rng = random.Random(8)
pairs = [rand_pair(rng, 50) for _ in range(5000)]
a = [None if rng.random() < 0.03 else p[0] for p in pairs]; b = [p[1] for p in pairs]
for m in oracle.MEASURES:
    ref, rv, ri = oracle.batch(m, a, b)
    v, valid, nulls, ints = _native.compute_host(m, pa.array(a, type=pa.string_view()), pa.array(b, type=pa.string_view()), debug=True)
    assert (valid == rv).all() and (v[rv].view(np.uint64) == ref[rv].view(np.uint64)).all() and (ints[rv] == ri[rv]).all(), m
    assert _native.last_overflow()[1] == int(rv.sum())
print("ok")
""" % (str(ROOT), str(ROOT / "polars-strsim_b200"), str(ROOT / "tests"))
    env = dict(os.environ, STRSIM_B200_FORCE_GENERIC="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_device_resident_api(native, oracle):
    torch = pytest.importorskip("torch")
    rng = random.Random(17)
    pairs = [rand_pair(rng, 24) for _ in range(50000)]
    a = [None if rng.random() < 0.02 else p[0] for p in pairs]
    b = [p[1] for p in pairs]
    ca, cb = native.DeviceColumn(sv(a)), native.DeviceColumn(sv(b))
    assert len(ca) == len(a) and ca.algorithmic_bytes > 16 * len(a)
    n = len(a)
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    val = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for measure in oracle.MEASURES:
        before = native.kernel_launches()
        native.compute_device(measure, ca, cb, out.data_ptr(), val.data_ptr(), 0, st)
        torch.cuda.synchronize()
        assert native.kernel_launches() > before
        ref, ref_valid, _ = oracle.batch(measure, a, b)
        got = out.cpu().numpy()
        bits = np.unpackbits(val.cpu().numpy().view(np.uint8), bitorder="little")[:n].astype(bool)
        assert (bits == ref_valid).all()
        assert (got[ref_valid].view(np.uint64) == ref[ref_valid].view(np.uint64)).all()


@pytest.mark.parametrize("n", [1, 511, 512, 513, 1536, 148 * 5 * 512 + 7])
def test_tiles_from_the_counter_cover_every_row_once(native, oracle, n):
    """short_kernel hands its tiles out through a counter that the last CTA of a launch resets: row counts
    around one tile (512 rows fused, 768 single measure), fewer tiles than CTAs, one tile more than the CTAs of
    a full grid; several launches back to back over the same overflow record -- single measures, the fused
    launch, and (medium rows mixed in) the 64-bit gather launch behind each of them -- must keep giving the
    oracle's rows, all of them, none twice (an output slot left from the run before would show)."""
    torch = pytest.importorskip("torch")
    rng = random.Random(1000 + n)
    base = [rand_pair(rng, 24) for _ in range(min(n, 4000))]
    medium = ("".join(rng.choice("abcdefgh") for _ in range(45)), "".join(rng.choice("abcdefgh") for _ in range(50)))
    a = [(medium[0] if i % 97 == 5 else base[i % len(base)][0]) for i in range(n)]
    b = [(medium[1] if i % 97 == 5 else base[i % len(base)][1]) for i in range(n)]
    ca, cb = native.DeviceColumn(sv(a)), native.DeviceColumn(sv(b))
    st = torch.cuda.current_stream().cuda_stream
    val = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
    refs = {m: oracle.batch(m, a[: len(base) * 2 + 100], b[: len(base) * 2 + 100])[0] for m in oracle.MEASURES}
    for rep in range(3):
        outs = [torch.full((n,), -1.0, dtype=torch.float64, device="cuda") for _ in oracle.MEASURES]
        native.compute_device_multi(list(oracle.MEASURES), ca, cb, [o.data_ptr() for o in outs], val.data_ptr(), None, st)
        single = torch.full((n,), -1.0, dtype=torch.float64, device="cuda")
        native.compute_device("jaro_winkler", ca, cb, single.data_ptr(), val.data_ptr(), 0, st)
        torch.cuda.synchronize()
        for m, o in zip(oracle.MEASURES, outs):
            got = o.cpu().numpy()
            k = len(refs[m])
            assert (got[:k].view(np.uint64) == refs[m].view(np.uint64)).all(), (m, rep)
            # the rows repeat with period len(base) except where the medium pair sits: same row, same bits
            if n > len(base):
                idx = np.arange(len(base), n)
                plain = (idx % 97 != 5) & ((idx % len(base)) % 97 != 5)
                assert (got[idx[plain]].view(np.uint64) == got[idx[plain] % len(base)].view(np.uint64)).all(), (m, rep)
            assert (got >= 0.0).all(), (m, rep, "a row was not written")
            if m == "jaro_winkler":
                assert (single.cpu().numpy().view(np.uint64) == got.view(np.uint64)).all(), rep


def test_multi_measure_single_upload(native, oracle):
    rng = random.Random(23)
    pairs = [rand_pair(rng, 40) for _ in range(20000)]
    a = [None if rng.random() < 0.04 else p[0] for p in pairs]
    b = [None if rng.random() < 0.04 else p[1] for p in pairs]
    outs, valid, nulls = native.compute_host_multi(list(oracle.MEASURES), sv(a), sv(b))
    for m, got in zip(oracle.MEASURES, outs):
        ref, ref_valid, _ = oracle.batch(m, a, b)
        assert (valid == ref_valid).all() and nulls == int((~ref_valid).sum())
        assert (got[ref_valid].view(np.uint64) == ref[ref_valid].view(np.uint64)).all(), m


def test_long_levenshtein_multiword(native, oracle):
    """Warp-cooperative multi-word Myers: block boundaries (64/128/2048 codepoints), 1-4 lanes'
    worth of blocks per lane, Unicode patterns, very unequal lengths."""
    rng = random.Random(99)
    a, b = [], []
    for la in (65, 127, 128, 129, 500, 2047, 2048, 2049, 4100, 6200):
        for lb in (0, 1, 64, 65, 300, 2048, 4099):
            alpha = rng.choice(["ab", "abcdefghijklmnopqrstuvwxyz ", "aé日\U0001f600z", "一二三四五六日本語"])
            x = "".join(rng.choice(alpha) for _ in range(la))
            y = "".join(rng.choice(alpha) for _ in range(lb))
            a += [x, x]
            b += [y, (x[: lb] + y[lb // 2:])[: max(lb, 1)]]
    # near-identical long pairs (edits sprinkled in)
    for _ in range(30):
        n = rng.randint(200, 4000)
        x = [rng.choice("abcdefghij ") for _ in range(n)]
        y = list(x)
        for _ in range(n // 10):
            p = rng.randrange(len(y))
            op = rng.randint(0, 2)
            if op == 0:
                y[p] = rng.choice("xyz")
            elif op == 1:
                y.insert(p, "q")
            elif len(y) > 1:
                del y[p]
        a.append("".join(x))
        b.append("".join(y))
    check(native, oracle, "levenshtein", a, b)
    assert native.last_overflow()[1] > 0


def test_polars_plugin_symbols_end_to_end(native, oracle):
    """Calls `_polars_plugin_<name>` the way polars-ffi does (fabricated SeriesExports, ownership of
    the inputs moves to the callee) and imports the returned Float64 series."""
    import ctypes

    import pyarrow as pa
    from polars_strsim._native import ArrowArray
    from test_abi import SeriesExport, make_series

    L = native.lib()
    rng = random.Random(61)
    pairs = [rand_pair(rng, 40) for _ in range(5000)]
    a = [None if rng.random() < 0.05 else p[0] for p in pairs]
    b = [p[1] for p in pairs]
    A = pa.chunked_array([pa.array(a[:777], type=pa.string_view()), pa.array(a[777:], type=pa.string_view())])
    B = pa.array(b, type=pa.string_view())
    for measure in oracle.MEASURES:
        released, keep = [], []
        inputs = (SeriesExport * 2)()
        inputs[0], arrs_a = make_series(A, released, keep)
        inputs[1], arrs_b = make_series(B, released, keep)
        ret = SeriesExport()
        ctx = ctypes.c_uint64(1)
        getattr(L, f"_polars_plugin_{measure}")(inputs, ctypes.c_size_t(2), None, ctypes.c_size_t(0),
                                                 ctypes.byref(ret), ctypes.byref(ctx))
        assert ret.private_data and ret.len == 1, native.lib().strsim_b200_last_error()
        assert len(released) == 2 and all(not x.release for x in arrs_a + arrs_b)
        # import like polars-ffi: move the ArrowArray out, borrow the field, then drop the SeriesExport
        moved = ArrowArray.from_address(ret.arrays[0])
        copy = ArrowArray()
        ctypes.memmove(ctypes.addressof(copy), ctypes.addressof(moved), ctypes.sizeof(ArrowArray))
        ret.release(ctypes.byref(ret))
        out = pa.Array._import_from_c(ctypes.addressof(copy), pa.float64())
        ref, ref_valid, _ = oracle.batch(measure, a, b)
        got_valid = np.array(out.is_valid())
        assert (got_valid == ref_valid).all()
        got = out.fill_null(0.0).to_numpy(zero_copy_only=False)
        assert (got[ref_valid] == ref[ref_valid]).all()
    # error path: shape mismatch leaves return_value untouched and stores the reference's message
    released, keep = [], []
    inputs = (SeriesExport * 2)()
    inputs[0], _ = make_series(pa.array(["a", "b", "c"], type=pa.string_view()), released, keep)
    inputs[1], _ = make_series(pa.array(["a", "b"], type=pa.string_view()), released, keep)
    ret = SeriesExport()
    L._polars_plugin_jaro(inputs, ctypes.c_size_t(2), None, ctypes.c_size_t(0), ctypes.byref(ret), None)
    assert not ret.private_data
    L._polars_plugin_get_last_error_message.restype = ctypes.c_char_p
    assert b"same length" in L._polars_plugin_get_last_error_message()


@pytest.mark.parametrize("alphabet", ["abcdefghijklmnopqrstuvwxyz", "ABCDEFxyz_`{|", "aB-c'D 0129.~\x01\x7f", "MiXeD case"])
def test_ascii_alphabet_blocks(native, oracle, alphabet):
    """ASCII columns whose bytes fit one 32- or 64-code-point block, or need all 7 bits: the
    5/6/7-plane instantiations of the register-resident path (row_ascii_reg.cuh)."""
    rng = random.Random(hash(alphabet) & 0xFFFF)
    a, b = [], []
    for _ in range(20000):
        x = "".join(rng.choice(alphabet) for _ in range(rng.randint(0, 32)))
        y = "".join(rng.choice(alphabet) for _ in range(rng.randint(0, 32))) if rng.random() < 0.4 else x
        if y is x and x:
            y = list(x)
            for _ in range(rng.randint(0, 3)):
                p = rng.randrange(len(y))
                y[p] = rng.choice(alphabet)
            if rng.random() < 0.5:
                y = y[1:]
            y = "".join(y)
        a.append(x)
        b.append(y)
    for measure in oracle.MEASURES:
        check(native, oracle, measure, a, b)


def test_long_levenshtein_tiers(native, oracle):
    """Pairs whose Peq does not fit the first-tier slab (many distinct codepoints x many blocks) and a
    pattern above 8192 codepoints (generic kernel) still give the exact distance."""
    rng = random.Random(2718)
    cjk = [chr(c) for c in range(0x4E00, 0x4E00 + 3000)]
    a, b = [], []
    for n in (3000, 5000):  # ~n distinct codepoints x n/64 blocks  >> 128 Ki words
        x = [rng.choice(cjk) for _ in range(n)]
        y = list(x)
        for _ in range(n // 20):
            y[rng.randrange(n)] = rng.choice(cjk)
        a.append("".join(x))
        b.append("".join(y[: n - 7]))
    x = "".join(rng.choice("abcdefgh") for _ in range(9000))
    y = "".join(rng.choice("abcdefgh") for _ in range(8500))
    a.append(x)
    b.append(y)
    a += ["short", "x" * 70]
    b += ["y" * 100, "x" * 69 + "z"]
    check(native, oracle, "levenshtein", a, b)


FUSED_SETS = [list(range(5)), [2, 4], [0, 1], [3, 4], [1, 2], [0, 3], [4, 2, 0], [1, 2, 3, 4]]


def check_multi(native, oracle, measure_ids, a, b, A=None, B=None):
    names = [oracle.MEASURES[i] for i in measure_ids]
    A = sv(a) if A is None else A
    B = sv(b) if B is None else B
    outs, valid, nulls, ints = native.compute_host_multi(names, A, B, debug=True)
    for m, got, gi in zip(names, outs, ints):
        ref, ref_valid, ref_ints = oracle.batch(m, a, b)
        assert (valid == ref_valid).all() and nulls == int((~ref_valid).sum()), (m, "null masks differ")
        bad = np.nonzero(valid & (got.view(np.uint64) != ref.view(np.uint64)))[0]
        assert bad.size == 0, (names, m, a[bad[0]], b[bad[0]], got[bad[0]], ref[bad[0]], gi[bad[0]], ref_ints[bad[0]])
        ibad = np.nonzero(valid & (gi != ref_ints).any(axis=1))[0]
        assert ibad.size == 0, (names, m, a[ibad[0]], b[ibad[0]], gi[ibad[0]], ref_ints[ibad[0]])


def test_fused_measures_golden_vectors(native, oracle):
    """ONE fused launch (short_kernel<MULTI_BASE+groups>) must reproduce every measure's golden bits."""
    fx = load_fixture()
    a, b = [r[1] for r in fx], [r[2] for r in fx]
    before = native.kernel_launches()
    check_multi(native, oracle, list(range(5)), a, b)
    # one fused launch + column-statistics and counter-readback launches (five single-measure launches
    # with their readbacks would make it 14 or more)
    assert native.kernel_launches() - before <= 9, "all five measures should come from one fused launch"
    demo_a = ["phillips", "phillips", "", "", None, None]
    demo_b = ["phillips", "philips", "phillips", "", "phillips", None]
    check_multi(native, oracle, list(range(5)), demo_a, demo_b)


@pytest.mark.parametrize("ids", FUSED_SETS)
def test_fused_measures_subsets(native, oracle, ids):
    """every group combination, on ASCII-only columns (bit-plane path), mixed scripts (register-compare
    path), with nulls, empties, single characters and rows that overflow to the 64-bit / long kernels"""
    rng = random.Random(1000 + sum(ids) * 7 + len(ids))
    # ASCII-only columns, lower-case block (5 planes)
    a, b = [], []
    for _ in range(30000):
        x = "".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(0, 30)))
        y = x
        if rng.random() < 0.8:
            y = list(x)
            for _ in range(rng.randint(0, 3)):
                p = rng.randint(0, len(y))
                op = rng.randint(0, 2)
                if op == 0 and y:
                    y[min(p, len(y) - 1)] = rng.choice("abcxyz")
                elif op == 1:
                    y.insert(p, rng.choice("abcxyz"))
                elif y:
                    del y[min(p, len(y) - 1)]
            y = "".join(y)
            if rng.random() < 0.2:
                y = "".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(0, 30)))
        a.append(x)
        b.append(y)
    a += ["a", "a", "", "x", "ab", "a" * 32, "a" * 33, "b" * 64, "c" * 65, "d" * 300]
    b += ["b", "a", "x", "", "ba", "a" * 31 + "b", "a" * 32, "b" * 63 + "c", "c" * 64, "e" * 200]
    check_multi(native, oracle, ids, a, b)
    # general ASCII (7 planes) with nulls
    a2 = [None if rng.random() < 0.05 else "".join(rng.choice("aB-c'D 019.~") for _ in range(rng.randint(0, 28))) for _ in range(8000)]
    b2 = [None if rng.random() < 0.05 else "".join(rng.choice("aB-c'D 019.~") for _ in range(rng.randint(0, 28))) for _ in range(8000)]
    check_multi(native, oracle, ids, a2, b2)
    # mixed scripts, lengths beyond 32 bytes included
    pairs = [rand_pair(rng, rng.choice([8, 20, 40, 90])) for _ in range(20000)]
    a3 = [None if rng.random() < 0.03 else p[0] for p in pairs]
    b3 = [None if rng.random() < 0.03 else p[1] for p in pairs]
    check_multi(native, oracle, ids, a3, b3)
    assert sum(native.last_overflow()) > 0


def test_fused_measures_chunks_broadcast_device(native, oracle):
    import pyarrow as pa
    torch = pytest.importorskip("torch")

    rng = random.Random(77)
    pairs = [rand_pair(rng, 24) for _ in range(40000)]
    a = [None if rng.random() < 0.05 else p[0] for p in pairs]
    b = [None if rng.random() < 0.05 else p[1] for p in pairs]
    # different chunkings + a sliced chunk
    A = pa.chunked_array([sv(a[:1000]), sv(a[1000:25000]), sv(["zz"] + a[25000:]).slice(1)])
    B = pa.chunked_array([sv(b[:17000]), sv(b[17000:])])
    check_multi(native, oracle, list(range(5)), a, b, A, B)
    # scalar broadcast, both orientations
    check_multi(native, oracle, [0, 2, 4], ["smith"] * len(b), b, sv(["smith"]), sv(b))
    check_multi(native, oracle, [1, 3], a, ["日本語"] * len(a), sv(a), sv(["日本語"]))
    # device-resident multi API
    ca, cb = native.DeviceColumn(sv(a)), native.DeviceColumn(sv(b))
    n = len(a)
    names = ["jaro_winkler", "sorensen_dice", "levenshtein"]
    outs = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in names]
    val = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
    before = native.kernel_launches()
    native.compute_device_multi(names, ca, cb, [o.data_ptr() for o in outs], val.data_ptr(), None,
                                torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    # fused kernel + validity kernel (+ follow-up launches per measure for the rows over 32 bytes)
    assert native.kernel_launches() - before >= 2
    bits = np.unpackbits(val.cpu().numpy().view(np.uint8), bitorder="little")[:n].astype(bool)
    for m, o in zip(names, outs):
        ref, ref_valid, _ = oracle.batch(m, a, b)
        got = o.cpu().numpy()
        assert (bits == ref_valid).all()
        assert (got[ref_valid].view(np.uint64) == ref[ref_valid].view(np.uint64)).all(), m
    with pytest.raises(native.StrsimError):
        native.compute_device_multi(["jaro", "jaro"], ca, cb, [outs[0].data_ptr(), outs[1].data_ptr()])


def test_fused_measures_fallback_cross_check(oracle):
    """fused call with STRSIM_B200_FORCE_GENERIC=1: every row leaves through the overflow lists and is
    finished measure by measure; and STRSIM_B200_NO_FUSE=1 must give the same bits as the fused path."""
    code = r"""
import sys, random, numpy as np
sys.path[:0] = [%r, %r, %r]
import pyarrow as pa
from polars_strsim import _native
from oracle import oracle
from test_oracle import rand_pair
rng = random.Random(9)
pairs = [rand_pair(rng, 50) for _ in range(4000)]
a = [None if rng.random() < 0.03 else p[0] for p in pairs]; b = [p[1] for p in pairs]
outs, valid, nulls, ints = _native.compute_host_multi(list(oracle.MEASURES), pa.array(a, type=pa.string_view()), pa.array(b, type=pa.string_view()), debug=True)
for m, v, gi in zip(oracle.MEASURES, outs, ints):
    ref, rv, ri = oracle.batch(m, a, b)
    assert (valid == rv).all() and (v[rv].view(np.uint64) == ref[rv].view(np.uint64)).all() and (gi[rv] == ri[rv]).all(), m
print("ok")
""" % (str(ROOT), str(ROOT / "polars-strsim_b200"), str(ROOT / "tests"))
    for knob in ("STRSIM_B200_FORCE_GENERIC", "STRSIM_B200_NO_FUSE"):
        env = dict(os.environ, **{knob: "1"})
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0 and "ok" in out.stdout, (knob, out.stderr[-2000:])


def test_progressive_upload_and_slice_redo(oracle):
    """Host calls send the data buffers progressively (a prefix per row slice).  Sequential columns need
    no second pass; a column whose early rows reference late bytes gets those slices recomputed."""
    code = r"""
import sys, random, numpy as np
sys.path[:0] = [%r, %r, %r]
import pyarrow as pa
from polars_strsim import _native
from oracle import oracle
rng = random.Random(3)
n = 60000
a = ["".join(rng.choice("abcdefghij") for _ in range(rng.randint(0, 30))) for _ in range(n)]
b = [x[: rng.randint(0, len(x))] + "".join(rng.choice("abcxyz") for _ in range(rng.randint(0, 8))) for x in a]
a[7] = None
A, B = pa.array(a, type=pa.string_view()), pa.array(b, type=pa.string_view())
def run(A, B, a, b, want_redo):
    outs, valid, nulls, ints = _native.compute_host_multi(list(oracle.MEASURES), A, B, debug=True)
    redo = _native.last_redo_slices()
    assert (redo > 0) == want_redo, redo
    for m, v, gi in zip(oracle.MEASURES, outs, ints):
        ref, rv, ri = oracle.batch(m, a, b)
        assert (valid == rv).all() and nulls == int((~rv).sum())
        assert (v[rv].view(np.uint64) == ref[rv].view(np.uint64)).all() and (gi[rv] == ri[rv]).all(), m
    v1, valid1, _ = _native.compute_host("jaro_winkler", A, B)
    ref, rv, _ = oracle.batch("jaro_winkler", a, b)
    assert (valid1 == rv).all() and (v1[rv].view(np.uint64) == ref[rv].view(np.uint64)).all()
run(A, B, a, b, False)
# a few rows swapped across the column: the views stay sequential almost everywhere (the sampled test
# passes) but rows near the start now reference bytes near the end of the data buffer
idx = np.arange(n)
for i, j in ((300, n - 5), (4500, n - 900), (n // 2 + 400, 30)):
    idx[i], idx[j] = idx[j], idx[i]
bufs = A.buffers()
views = np.frombuffer(bufs[1], dtype=np.int32).reshape(-1, 4)
valid = np.ones(n, dtype=bool); valid[7] = False
vbits = np.packbits(valid[idx], bitorder="little")
A2 = pa.Array.from_buffers(pa.string_view(), n, [pa.py_buffer(vbits.tobytes()), pa.py_buffer(np.ascontiguousarray(views[idx]).tobytes())] + bufs[2:])
a2 = [a[i] for i in idx]
assert A2.to_pylist() == a2
run(A2, B, a2, b, True)
print("ok")
""" % (str(ROOT), str(ROOT / "polars-strsim_b200"), str(ROOT / "tests"))
    env = dict(os.environ, STRSIM_B200_SLICE_ROWS="4096")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-3000:]


def test_plugin_calls_reuse_columns_already_in_hbm(native, oracle):
    """The five README expressions are five plugin calls over the same two columns: the first call
    uploads (and keeps) them, the other four must find them in HBM -- and the arrays the library took
    ownership of stay alive until it lets go of them."""
    import ctypes
    import gc

    import pyarrow as pa
    from polars_strsim._native import ArrowArray
    from test_abi import SeriesExport, make_series

    L = native.lib()
    L.strsim_b200_cache_stats.argtypes = [ctypes.POINTER(ctypes.c_int64)]
    L.strsim_b200_cache_stats.restype = None
    L.strsim_b200_cache_clear.restype = None

    def stats():
        out = (ctypes.c_int64 * 4)()
        L.strsim_b200_cache_stats(out)
        return list(out)

    def call(measure, A, B):
        released, keep = [], []
        inputs = (SeriesExport * 2)()
        inputs[0], arrs_a = make_series(A, released, keep)
        inputs[1], arrs_b = make_series(B, released, keep)
        ret = SeriesExport()
        getattr(L, f"_polars_plugin_{measure}")(inputs, ctypes.c_size_t(2), None, ctypes.c_size_t(0),
                                                 ctypes.byref(ret), None)
        assert ret.private_data and ret.len == 1, L.strsim_b200_last_error()
        assert len(released) == 2 and all(not x.release for x in arrs_a + arrs_b)
        moved = ArrowArray.from_address(ret.arrays[0])
        copy = ArrowArray()
        ctypes.memmove(ctypes.addressof(copy), ctypes.addressof(moved), ctypes.sizeof(ArrowArray))
        ret.release(ctypes.byref(ret))
        return pa.Array._import_from_c(ctypes.addressof(copy), pa.float64())

    def check_out(measure, out, a, b):
        ref, ref_valid, _ = oracle.batch(measure, a, b)
        assert (np.array(out.is_valid()) == ref_valid).all()
        got = out.fill_null(0.0).to_numpy(zero_copy_only=False)
        assert (got[ref_valid].view(np.uint64) == ref[ref_valid].view(np.uint64)).all(), measure

    L.strsim_b200_cache_clear()
    L.strsim_b200_speculation(0)  # this test counts cache hits call by call: no results computed ahead
    rng = random.Random(5)
    n = 120000
    a = [None if rng.random() < 0.02 else "".join(rng.choice("abcdefghij") for _ in range(rng.randint(0, 28))) for _ in range(n)]
    b = [x if x is None or rng.random() < 0.3 else x[: rng.randint(0, len(x))] + rng.choice(["", "xy", "k"]) for x in a]
    A = pa.chunked_array([sv(a[:50000]), sv(a[50000:])])
    B = sv(b)
    h0, m0, _, _ = stats()
    for k, measure in enumerate(oracle.MEASURES):
        check_out(measure, call(measure, A, B), a, b)
        h, m, cols, nbytes = stats()
        assert (h - h0, m - m0) == (2 * k, 2), (measure, h - h0, m - m0)
        assert cols == 2 and nbytes > 16 * n
    # a scalar operand is never cached, the column operand still hits
    check_out("jaro", call("jaro", A, sv(["abcdef"])), a, ["abcdef"] * n)
    assert stats()[0] - h0 == 2 * 4 + 1
    # the library holds the arrays: dropping ours must not invalidate what the cache identifies by address
    del A, B
    gc.collect()
    a2 = [x[::-1] if x else x for x in a]
    A2, B2 = sv(a2), sv(b)
    check_out("levenshtein", call("levenshtein", A2, B2), a2, b)  # new buffers: misses, correct results
    assert stats()[2] == 4
    L.strsim_b200_cache_clear()
    assert stats()[2:] == [0, 0]
    check_out("sorensen_dice", call("sorensen_dice", A2, B2), a2, b)
    L.strsim_b200_speculation(1)


def test_concurrent_plugin_calls_share_one_upload(native, oracle):
    """Polars evaluates the expressions of one `with_columns` on several threads: the five README calls may
    arrive together.  Exactly ONE of them uploads each column (two misses in all); the others wait for that
    upload instead of repeating it, and every result is bit-exact.  Results of this size land in pinned
    buffers of the plugin's pool once the pool has grown (second round)."""
    import ctypes
    import threading
    import time

    sys.path.insert(0, str(ROOT))
    from bench_support import plugin_driver, workloads

    L = native.lib()
    L.strsim_b200_cache_stats.argtypes = [ctypes.POINTER(ctypes.c_int64)]
    L.strsim_b200_cache_stats.restype = None

    def stats():
        out = (ctypes.c_int64 * 4)()
        L.strsim_b200_cache_stats(out)
        return list(out)

    n = 1_500_000
    A, B = workloads.make_pairs(2, n)
    a, b = A.to_pylist(), B.to_pylist()
    refs = {m: oracle.batch(m, a, b)[0] for m in oracle.MEASURES}
    L.strsim_b200_speculation(0)  # every call computes its own measure here (test_companion_measures covers the rest)
    for round_ in range(2):
        plugin_driver.cache_clear()
        h0, m0, _, _ = stats()
        got, errors = {}, []

        def work(measure):
            try:
                r = plugin_driver.call(measure, A, B)
                got[measure] = np.concatenate(r.values()).copy()
                r.release()
            except Exception as exc:  # noqa: BLE001
                errors.append((measure, exc))

        threads = [threading.Thread(target=work, args=(m,)) for m in oracle.MEASURES]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errors, errors
        h, m, cols, _ = stats()
        assert m - m0 == 2 and cols == 2, ("every column must be uploaded exactly once", h - h0, m - m0, cols)
        assert h - h0 == 8
        for measure in oracle.MEASURES:
            assert (got[measure].view(np.uint64) == refs[measure].view(np.uint64)).all(), measure
        time.sleep(0.5)  # lets the background thread pin result buffers for the second round
    plugin_driver.cache_clear()
    L.strsim_b200_speculation(1)


def test_companion_measures_are_computed_with_the_upload(native, oracle):
    """The README query is five plugin calls over the same two columns (README.md:47-51).  Once the plugin has
    SEEN such a query, the call that uploads the columns of the next one computes the other four measures in
    the same fused pass and downloads them behind the upload; the later calls take their result ready-made
    (no kernel launch at all) -- bit-exact, with the null mask, sequentially and from five threads at once.
    Nothing is computed ahead for a process that asks for one measure only."""
    import ctypes
    import threading

    sys.path.insert(0, str(ROOT))
    from bench_support import plugin_driver, workloads

    L = native.lib()
    L.strsim_b200_speculation_stats.argtypes = [ctypes.POINTER(ctypes.c_int64)]
    L.strsim_b200_speculation_stats.restype = None

    def spec():
        out = (ctypes.c_int64 * 3)()
        L.strsim_b200_speculation_stats(out)
        return list(out)

    n = 300_000
    plugin_driver.cache_clear()
    L.strsim_b200_speculation(1)  # also forgets what earlier tests taught it
    queries = [workloads.make_pairs(3, n, row_base=q * n) for q in range(4)]  # nulls, mixed scripts

    def run_query(A, B, measures, threads=False):
        got = {}

        def work(m):
            r = plugin_driver.call(m, A, B)
            got[m] = (np.concatenate(r.values()).copy(), r.validity(), r.null_count)
            r.release()
        if threads:
            ts = [threading.Thread(target=work, args=(m,)) for m in measures]
            [t.start() for t in ts]
            [t.join() for t in ts]
        else:
            for m in measures:
                work(m)
        a, b = A.to_pylist(), B.to_pylist()
        for m in measures:
            ref, rv, _ = oracle.batch(m, a, b)
            vals, valid, nulls = got[m]
            assert (valid == rv).all() and nulls == int((~rv).sum()), m
            assert (vals[rv].view(np.uint64) == ref[rv].view(np.uint64)).all(), m

    served0 = spec()[0]
    run_query(*queries[0], oracle.MEASURES)              # the first query teaches the companions
    assert spec()[0] == served0 and spec()[1] == 0
    launches = native.kernel_launches()
    run_query(*queries[1], oracle.MEASURES)              # the second one gets four results ready-made
    assert spec()[0] - served0 == 4 and spec()[1] == 0
    assert spec()[2] == 0b11111
    run_query(*queries[2], oracle.MEASURES, threads=True)  # all five calls at once
    assert spec()[0] - served0 == 8
    # a query that asks for one measure only: what was computed ahead for it waits, then expires -- and the
    # query after it computes nothing ahead any more
    run_query(*queries[3], ["jaro"])
    assert spec()[1] == 4
    run_query(*queries[0], ["jaro"])
    assert spec()[2] == 0b00010
    waiting = spec()[1]
    run_query(*queries[1], ["jaro"])
    assert spec()[1] <= waiting  # nothing new computed ahead (what waited may have expired meanwhile)
    plugin_driver.cache_clear()
    assert spec()[1] == 0
    L.strsim_b200_speculation(1)


def test_dictionary_encoded_columns(native, oracle):
    """Arrow dictionary / Polars Categorical inputs (the reference rejects them at
    `.str()?`, strsim.rs:46-47): only the indices and the dictionary are uploaded, the rows' views are
    materialised on the device, and the results equal those of the decoded String columns bit for bit --
    through the Arrow entry point (borrowed inputs) and through the plugin symbols (owned inputs, cached),
    with every index width, null indices, null dictionary entries, sliced and multi-chunk arrays, a plain
    String column on the other side, long entries (> 64 bytes) and a literal."""
    import pyarrow as pa

    sys.path.insert(0, str(ROOT))
    from bench_support import plugin_driver
    from polars_strsim import arrow as strsim_arrow

    rng = random.Random(77)
    words = ["", "a", "smith", "smyth", "josé maría", "日本語", "phillips", "philips", None,
             "a-rather-long-surname-over-32-bytes-long", "x" * 70 + "yz", "naïve", "Ünal-Çelik"]
    words += ["".join(rng.choice("abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(1, 24))) for _ in range(200)]
    n = 70_000

    def encoded(index_type, seed, values_type=pa.string_view()):
        r = random.Random(seed)
        idx = [None if r.random() < 0.03 else r.randrange(len(words)) for _ in range(n)]
        cap = {pa.int8(): 127, pa.uint8(): 255}.get(index_type)
        if cap is not None:
            idx = [i if i is None or i <= cap else i % cap for i in idx]
        arr = pa.DictionaryArray.from_arrays(pa.array(idx, type=index_type), pa.array(words, type=values_type))
        return arr, [None if i is None else words[i] for i in idx]

    def check_values(measure, got, a, b):
        ref, rv, _ = oracle.batch(measure, a, b)
        assert (np.asarray(got.is_valid()) == rv).all(), measure
        vals = got.fill_null(0.0).to_numpy(zero_copy_only=False)
        assert (vals[rv].view(np.uint64) == ref[rv].view(np.uint64)).all(), measure

    B_plain_list = [None if rng.random() < 0.02 else rng.choice([w for w in words if w is not None]) for _ in range(n)]
    B_plain = sv(B_plain_list)
    for k, (itype, vtype) in enumerate([(pa.int32(), pa.string_view()), (pa.uint32(), pa.large_string()),
                                        (pa.int8(), pa.string()), (pa.uint16(), pa.string_view()),
                                        (pa.int64(), pa.string_view())]):
        A, a = encoded(itype, 100 + k, vtype)
        Bd, b = encoded(pa.uint32(), 200 + k)
        measure = oracle.MEASURES[k % 5]
        check_values(measure, strsim_arrow.compute(measure, A, Bd), a, b)              # both encoded
        check_values(measure, strsim_arrow.compute(measure, A, B_plain), a, B_plain_list)  # one plain String column
        check_values(measure, strsim_arrow.compute(measure, A.slice(17, 50_001), Bd.slice(300, 50_001)),
                     a[17:50_018], b[300:50_301])                                     # ArrowArray.offset on indices + validity
    # multi-chunk, through the plugin symbols: the second call finds the materialised columns in HBM
    A, a = encoded(pa.uint32(), 7)
    Bd, b = encoded(pa.uint32(), 8)
    CA = pa.chunked_array([A.slice(0, 30_000), A.slice(30_000)])
    plugin_driver.cache_clear()
    for measure in oracle.MEASURES:
        r = plugin_driver.call(measure, CA, Bd)
        got = np.concatenate(r.values())
        valid = r.validity()
        ref, rv, _ = oracle.batch(measure, a, b)
        assert (valid == rv).all() and r.null_count == int((~rv).sum()), measure
        assert (got[rv].view(np.uint64) == ref[rv].view(np.uint64)).all(), measure
        r.release()
    plugin_driver.cache_clear()
    # a literal against an encoded column; a null literal still fails the call
    check_values("jaro_winkler", strsim_arrow.compute("jaro_winkler", A, "smith"), a, ["smith"] * n)
    with pytest.raises(native.StrsimError, match="literal operand is null"):
        strsim_arrow.compute("jaro", A, None)
    # a non-String dictionary is a dtype error like any other non-String column
    bad = pa.DictionaryArray.from_arrays(pa.array([0, 1], type=pa.int32()), pa.array([1.5, 2.5]))
    with pytest.raises(native.StrsimError, match="expected `String`"):
        strsim_arrow.compute("jaro", bad, bad)


def test_pageable_inputs_and_outputs_go_through_the_pinned_rings(oracle):
    """Ordinary (pageable) Arrow buffers in, ordinary numpy arrays out: uploads are staged through the
    pinned ring by the copy threads, one row slice ahead of the kernels; downloads likewise."""
    code = r"""
import sys, random, numpy as np
sys.path[:0] = [%r, %r, %r]
import pyarrow as pa
from polars_strsim import _native
from oracle import oracle
rng = random.Random(12)
n = 450000
words = ["".join(rng.choice("abcdefghijklmnop") for _ in range(rng.randint(0, 30))) for _ in range(5000)]
a = [rng.choice(words) for _ in range(n)]
b = [x[: rng.randint(0, len(x))] + rng.choice(["", "q", "zz"]) if rng.random() < 0.7 else rng.choice(words) for x in a]
b[11] = None
A, B = pa.array(a, type=pa.string_view()), pa.array(b, type=pa.string_view())
assert A.buffers()[1].size > (4 << 20)
outs, valid, nulls = _native.compute_host_multi(list(oracle.MEASURES), A, B)
assert _native.last_redo_slices() == 0
for m, v in zip(oracle.MEASURES, outs):
    ref, rv, _ = oracle.batch(m, a, b)
    assert (valid == rv).all() and nulls == 1
    assert (v[rv].view(np.uint64) == ref[rv].view(np.uint64)).all(), m
v1, valid1, _ = _native.compute_host("levenshtein", A, B)
ref, rv, _ = oracle.batch("levenshtein", a, b)
assert (v1[rv].view(np.uint64) == ref[rv].view(np.uint64)).all()
print("ok")
""" % (str(ROOT), str(ROOT / "polars-strsim_b200"), str(ROOT / "tests"))
    env = dict(os.environ, STRSIM_B200_SLICE_ROWS="100000")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-3000:]


def _sample_rows(col, idx):
    import pyarrow as pa

    flat = col.combine_chunks() if isinstance(col, pa.ChunkedArray) else col
    return flat.cast(pa.large_string()).take(pa.array(idx)).to_pylist()  # take() has no string_view kernel


def _mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


@pytest.mark.parametrize("config,rows,row_base", [(2, 10_000_000, 0), (3, 100_000_000, 0), (7, 2_000_000, 0),
                                                   (8, 2_000_000, 0), (5, 10_000_000, 3 * 125_000_000)])
def test_full_size_workloads_by_properties(native, oracle, config, rows, row_base):
    """BASELINE configs C2 and C3 at their FULL size (10 M / 100 M rows), 10 M rows out of the middle of C5's
    1 B (the shard rank 3 of 8 owns: generator seed 0xC5, rows 375 M..385 M, Jaro-Winkler + Sorensen-Dice),
    2 M medium ASCII strings of 20-60 characters (config 7: three quarters of the rows run the 64-bit plane
    kernel) and 2 M mixed-case names (config 8: the 7-plane kernel) -- where the oracle cannot check every
    row in seconds: size-independent properties over ALL rows, plus the oracle bit for bit (values and
    integer intermediates) on a seeded sample of 100k rows.  Properties: values in [0, 1]; null mask = AND
    of the input masks; the symmetric measures (Levenshtein, Jaccard, Sorensen-Dice: edit distance and
    multiset intersection do not depend on the argument order, strsim.rs:146-160,297-306) give identical
    bits with the columns swapped -- which tables the OTHER string and streams the other one through every
    kernel; byte-equal rows score exactly 1.0 and only they do for Levenshtein (d = 0 iff equal); and, where
    the debug records of all rows fit the host (up to 20 M rows), inter and the lengths in the records
    satisfy x1_jaccard + inter = x1_dice, Jaro and Jaro-Winkler share m and t."""
    import pyarrow as pa
    import pyarrow.compute  # noqa: F401

    sys.path.insert(0, str(ROOT))
    from bench_support import workloads

    if rows > 20_000_000 and _mem_available_gb() < 40:
        pytest.skip("needs about 30 GB of host memory for the columns and two sets of results")
    A, B = workloads.make_pairs(config, rows, row_base=row_base, uneven_b=(config == 3))
    names = ["jaro_winkler", "sorensen_dice"] if config == 5 else list(oracle.MEASURES)
    with_ints = rows <= 20_000_000
    if with_ints:
        outs, valid, nulls, ints = native.compute_host_multi(names, A, B, debug=True)
    else:
        outs, valid, nulls = native.compute_host_multi(names, A, B)
        ints = None
    swapped, valid_s, nulls_s = native.compute_host_multi(names, B, A)
    n = rows
    assert all(len(o) == n for o in outs)

    def flat(col):
        return col.combine_chunks() if isinstance(col, pa.ChunkedArray) else col

    def mask(col):
        return np.asarray(flat(col).is_valid())
    expect_valid = mask(A) & mask(B)
    assert (valid == expect_valid).all() and (valid_s == expect_valid).all()
    assert nulls == int((~expect_valid).sum()) == nulls_s
    by = dict(zip(names, outs))
    by_s = dict(zip(names, swapped))
    for m in names:
        v = by[m][valid]
        assert np.isfinite(v).all() and (v >= 0.0).all() and (v <= 1.0).all(), m
    for m in ("levenshtein", "jaccard", "sorensen_dice"):
        if m in by:
            assert (by[m][valid].view(np.uint64) == by_s[m][valid].view(np.uint64)).all(), (m, "not symmetric")
    del swapped, by_s
    # byte-equal rows, found on the host from the Arrow buffers themselves
    equal_rows = valid & np.asarray(pa.compute.equal(flat(A).cast(pa.large_string()), flat(B).cast(pa.large_string()))
                                    .fill_null(False))
    for m in names:
        assert (by[m][equal_rows] == 1.0).all(), m
    if "levenshtein" in by:
        assert ((by["levenshtein"] == 1.0) & valid).sum() == equal_rows.sum()
    if "jaro" in by:
        assert (by["jaro_winkler"][valid] >= by["jaro"][valid]).all()
    if with_ints:
        iby = dict(zip(names, ints))
        for m in names:
            assert ((iby[m][:, 0] == 1) & valid == equal_rows).all(), (m, "F_EQUAL record differs from the host's compare")
        if "jaccard" in iby:
            gen = valid & (iby["jaccard"][:, 0] == 0)
            assert (iby["jaccard"][gen, 3] == iby["sorensen_dice"][gen, 3]).all()                      # same intersection
            assert (iby["jaccard"][gen, 4] + iby["jaccard"][gen, 3] == iby["sorensen_dice"][gen, 4]).all()  # uni + inter = la + lb
        if "jaro" in iby:
            assert (iby["jaro"][valid, 3] == iby["jaro_winkler"][valid, 3]).all()                      # same match count
            assert (iby["jaro"][valid, 4] == iby["jaro_winkler"][valid, 4]).all()                      # same transpositions
    # the oracle on a seeded sample: values bit for bit against the full-size run; the integer intermediates
    # from the full run's records, or (100 M rows) from a second call over the sampled rows alone
    rng = np.random.default_rng(config)
    idx = np.sort(rng.choice(n, size=100_000, replace=False))
    a, b = _sample_rows(A, idx), _sample_rows(B, idx)
    if not with_ints:
        s_outs, s_valid, _, s_ints = native.compute_host_multi(names, sv(a), sv(b), debug=True)
        for m, so in zip(names, s_outs):
            assert (so[s_valid].view(np.uint64) == by[m][idx][s_valid].view(np.uint64)).all(), (m, "sample call != full run")
        iby_sample = dict(zip(names, s_ints))
    for m in names:
        ref, ref_valid, ref_ints = oracle.batch(m, a, b)
        got, gv = by[m][idx], valid[idx]
        gi = iby[m][idx] if with_ints else iby_sample[m]
        assert (gv == ref_valid).all(), m
        bad = np.nonzero(gv & (got.view(np.uint64) != ref.view(np.uint64)))[0]
        assert bad.size == 0, (m, a[bad[0]], b[bad[0]], got[bad[0]], ref[bad[0]])
        ibad = np.nonzero(gv & (gi != ref_ints).any(axis=1))[0]
        assert ibad.size == 0, (m, a[ibad[0]], b[ibad[0]], gi[ibad[0]], ref_ints[ibad[0]])


def test_long_levenshtein_c4_by_properties(native, oracle):
    """C4 (long-text pairs, 200-4000 codepoints, Levenshtein) at its FULL size of 1 M pairs (4.4e12 DP cells):
    symmetry under swapped columns over ALL rows (the warp-cooperative kernel then tables the same shorter
    string but reads it from the other column), the bounds |la - lb| <= d <= max(la, lb) from the kernel's
    own record, and the oracle (two-row DP, strsim.rs:146-159) bit for bit on a seeded sample of 400 rows."""
    sys.path.insert(0, str(ROOT))
    from bench_support import workloads

    n = 1_000_000
    if _mem_available_gb() < 24:
        pytest.skip("needs about 12 GB of host memory for the two 2.3 GB columns and their copies")
    A, B = workloads.make_pairs(4, n)
    vals, valid, nulls, ints = native.compute_host("levenshtein", A, B, debug=True)
    swapped, valid_s, _, ints_s = native.compute_host("levenshtein", B, A, debug=True)
    assert valid.all() and valid_s.all() and nulls == 0
    assert (vals.view(np.uint64) == swapped.view(np.uint64)).all()
    assert (ints[:, 3] == ints_s[:, 3]).all() and (ints[:, 1] == ints_s[:, 2]).all()
    gen = ints[:, 0] == 0
    la, lb, d = ints[gen, 1].astype(np.int64), ints[gen, 2].astype(np.int64), ints[gen, 3].astype(np.int64)
    assert (d >= np.abs(la - lb)).all() and (d <= np.maximum(la, lb)).all() and (d > 0).all()
    assert ((vals >= 0.0) & (vals <= 1.0)).all()
    rng = np.random.default_rng(4)
    idx = np.sort(rng.choice(n, size=400, replace=False))
    a, b = _sample_rows(A, idx), _sample_rows(B, idx)
    ref, ref_valid, ref_ints = oracle.batch("levenshtein", a, b)
    assert (vals[idx].view(np.uint64) == ref.view(np.uint64)).all()
    assert (ints[idx] == ref_ints).all()


def test_concurrent_calls_from_engine_threads(native, oracle):
    """`is_elementwise=True` (polars_strsim/__init__.py:15) lets the engine call the plugin per chunk from
    any of its threads at the same time: the exports must be re-entrant.  Eight threads call the C ABI
    concurrently (ctypes drops the GIL), each with its own rows -- short ASCII, mixed scripts, nulls, rows
    for the 64-bit and the long kernels -- single measures and fused sets, three rounds each; every result
    is compared with the oracle bit for bit."""
    import threading

    rng = random.Random(77)
    jobs = []
    for t in range(8):
        a, b = [], []
        for _ in range(3000):
            x, y = rand_pair(rng)
            a.append(x)
            b.append(y)
        for k in range(40):  # some rows for the 64-bit and the long kernels, some nulls
            s = "".join(rng.choice("abcdefgh ") for _ in range(rng.randrange(30, 400)))
            a.append(s)
            b.append(s[: len(s) // 2] + "x" + s[len(s) // 2 + (k % 3):])
        a[5] = None
        b[t + 10] = None
        names = [oracle.MEASURES[(t + i) % 5] for i in range(1 + t % 3)]
        jobs.append((a, b, names))
    refs = [[oracle.batch(m, a, b) for m in names] for a, b, names in jobs]
    errors = []

    def work(idx):
        a, b, names = jobs[idx]
        try:
            A, B = sv(a), sv(b)
            for _ in range(3):
                if len(names) == 1:
                    vals, valid, nulls, ints = native.compute_host(names[0], A, B, debug=True)
                    outs, all_ints = [vals], [ints]
                else:
                    outs, valid, nulls, all_ints = native.compute_host_multi(names, A, B, debug=True)
                for (ref, ref_valid, ref_ints), got, gi in zip(refs[idx], outs, all_ints):
                    assert (valid == ref_valid).all()
                    assert (got.view(np.uint64)[valid] == ref.view(np.uint64)[valid]).all()
                    assert (gi[valid] == ref_ints[valid]).all()
        except BaseException as exc:  # noqa: BLE001 -- reported by the main thread
            errors.append((idx, repr(exc)))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


def test_edge_rows_of_the_plane_kernels(native, oracle):
    """Rows that sit on the special cases of the plane kernels' sort key and sources: NUL and control
    characters (a zero byte equals the zero padding, so only the lengths may tell them apart), columns in
    which every row is byte-equal (all rows in the 'equal' bucket), an empty string against a non-empty
    one in ASCII and in Latin-1 columns (key of the empty streamed string next to the equal bucket),
    strings of exactly 32 one-byte characters, and prefixes that differ only past the shorter string."""
    cases = [
        (["a\x00b", "\x00", "\x00\x00", "ab\x00", "\x00abc", "abc", ""], ["a\x00c", "", "\x00", "ab", "\x00abd", "abc\x00", "\x00"]),
        (["same", "same", "", "x" * 32, "yy"] * 300, ["same", "same", "", "x" * 32, "yy"] * 300),
        (["", "abc", "", "é", "", "ab"] * 50, ["abc", "", "é", "", "", "ab"] * 50),
        (["é" * 16, "a" * 32, "ab" * 16, "é" * 16 + "a", "z" * 31 + "é"], ["é" * 15 + "e", "a" * 31 + "b", "ba" * 16, "é" * 16, "z" * 32]),
        (["abcd", "abc", "abcde", "ab"], ["abc", "abcd", "abcdx", "abxx"]),
    ]
    for a, b in cases:
        for m in oracle.MEASURES:
            check(native, oracle, m, a, b)
        check_multi(native, oracle, list(range(5)), a, b)
        check_multi(native, oracle, [2, 4], a, b)
