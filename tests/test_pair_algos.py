"""Host-side check of the product's per-pair templates (csrc/pair_algos.cuh, csrc/row_short.cuh).

The same templates the CUDA kernels instantiate are compiled for the host by
tests/host_harness/pair_algos_host.cpp and compared with the oracle: integer intermediates and f64
bits must be identical.  CPU only -- this validates the bit-parallel arithmetic before any GPU time.
"""
import ctypes
import random
import subprocess
from pathlib import Path

import numpy as np
import pytest

from test_oracle import load_fixture, rand_pair

ROOT = Path(__file__).resolve().parent.parent
HARNESS = ROOT / "tests" / "host_harness"


@pytest.fixture(scope="module")
def algos():
    so = HARNESS / "libpair_algos_host.so"
    srcs = [HARNESS / "pair_algos_host.cpp"] + sorted((ROOT / "polars-strsim_b200/csrc").glob("*.cuh"))
    if not so.exists() or so.stat().st_mtime < max(s.stat().st_mtime for s in srcs):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                        f"-I{ROOT / 'polars-strsim_b200/csrc'}", "-o", str(so), str(srcs[0])], check=True)
    L = ctypes.CDLL(str(so))
    L.algos_batch.restype = ctypes.c_int
    L.algos_batch.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64] + [ctypes.c_void_p] * 4 + \
        [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return L


def run_batch(L, oracle, measure, bits, a, b, force_unicode=False):
    from oracle.oracle import _pack, MEASURE_ID

    ad, ao, _ = _pack(a)
    bd, bo, _ = _pack(b)
    n = len(a)
    ints = np.zeros((n, 6), dtype=np.int32)
    vals = np.zeros(n, dtype=np.float64)
    bad = L.algos_batch(MEASURE_ID[measure], bits, n, ad.ctypes.data, ao.ctypes.data, bd.ctypes.data,
                        bo.ctypes.data, int(force_unicode), ints.ctypes.data, vals.ctypes.data)
    assert bad == 0, "PM table not restored to zero / length over capacity"
    ref, _, ref_ints = oracle.batch(measure, a, b)
    mism = np.nonzero(vals.view(np.uint64) != ref.view(np.uint64))[0]
    assert mism.size == 0, (measure, bits, a[mism[0]], b[mism[0]], vals[mism[0]], ref[mism[0]])
    imism = np.nonzero((ints != ref_ints).any(axis=1))[0]
    assert imism.size == 0, (measure, bits, a[imism[0]], b[imism[0]], ints[imism[0]], ref_ints[imism[0]])


@pytest.mark.parametrize("bits", [32, 64])
@pytest.mark.parametrize("force_unicode", [False, True])
def test_reference_vectors(algos, oracle, bits, force_unicode):
    fx = load_fixture()
    for measure in oracle.MEASURES:
        rows = [r for r in fx if r[0] == measure]
        run_batch(algos, oracle, measure, bits, [r[1] for r in rows], [r[2] for r in rows], force_unicode)


@pytest.mark.parametrize("bits", [32, 64])
def test_random_pairs(algos, oracle, bits):
    rng = random.Random(99 + bits)
    a, b = [], []
    while len(a) < 40000:
        x, y = rand_pair(rng, bits)
        if len(x.encode()) <= bits and len(y.encode()) <= bits:
            a.append(x)
            b.append(y)
    for measure in oracle.MEASURES:
        run_batch(algos, oracle, measure, bits, a, b)
        run_batch(algos, oracle, measure, bits, a[:8000], b[:8000], force_unicode=True)


@pytest.mark.parametrize("bits", [32, 64])
def test_boundary_lengths(algos, oracle, bits):
    """exactly-full words, empty sides, single characters, 1-4 byte UTF-8"""
    rng = random.Random(5)
    a, b = [], []
    for la in (0, 1, 2, bits - 1, bits):
        for lb in (0, 1, 2, bits - 1, bits):
            for _ in range(40):
                a.append("".join(rng.choice("abc") for _ in range(la)))
                b.append("".join(rng.choice("abc") for _ in range(lb)))
    for ch in ("é", "日", "\U0001f600"):
        w = len(ch.encode())
        for la in (1, 2, bits // w):
            for lb in (1, 2, bits // w):
                a.append(ch * la)
                b.append((ch * lb)[:-1] + "x" if lb > 1 else ch)
                a.append(ch * la)
                b.append("x" * lb)
    for measure in oracle.MEASURES:
        run_batch(algos, oracle, measure, bits, a, b)


def test_multiword_myers_block(algos, oracle):
    """The 64-cell block step of the long-string kernel, chained over blocks on the host."""
    algos.algos_myers_multiword.restype = ctypes.c_int
    algos.algos_myers_multiword.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    rng = random.Random(4242)
    cases = [rand_pair(rng, L) for L in (5, 63, 64, 65, 127, 128, 129, 300, 700) for _ in range(25)]
    cases += [("a" * 200, "a" * 199 + "b"), ("", "x" * 100), ("ab" * 100, "ba" * 100)]
    for x, y in cases:
        ax = np.array([ord(c) for c in x], dtype=np.uint32)
        ay = np.array([ord(c) for c in y], dtype=np.uint32)
        d = algos.algos_myers_multiword(ax.ctypes.data, len(x), ay.ctypes.data, len(y))
        if x == y:
            assert d == 0
        else:
            assert d == oracle.pair("levenshtein", x, y)[1][3], (len(x), len(y))


@pytest.mark.parametrize("nbits,alphabet", [(5, "abcdefghijklmnopqrstuvwxyz"), (5, "ABCXYZ@["),
                                            (6, "abcxyzABCXYZ_`{"), (7, "ab yz-'09AZ~\x01\x7f!"),
                                            (107, "ab yz-'09AZ~\x01\x7f!"), (107, "abcdefghij ")])
def test_register_resident_ascii_path(algos, oracle, nbits, alphabet):
    """row_ascii_reg.cuh: strings in registers, position masks from bit planes (no table)."""
    from oracle.oracle import _pack, MEASURE_ID

    algos.algos_batch_reg.restype = ctypes.c_int
    algos.algos_batch_reg.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64] + [ctypes.c_void_p] * 6
    rng = random.Random(1000 + nbits)
    cap = 64 if nbits >= 100 else 32  # nbits 107 = 64-bit masks (strings of up to 64 characters), 7 planes
    a, b = [], []
    for _ in range(20000):
        x = "".join(rng.choice(alphabet) for _ in range(rng.randint(0, cap)))
        if rng.random() < 0.6:
            y = list(x)
            for _ in range(rng.randint(0, 3)):
                op, pos = rng.randint(0, 3), rng.randint(0, len(y))
                if op == 0 and y:
                    y[min(pos, len(y) - 1)] = rng.choice(alphabet)
                elif op == 1 and len(y) < cap:
                    y.insert(pos, rng.choice(alphabet))
                elif op == 2 and y:
                    del y[min(pos, len(y) - 1)]
                elif op == 3 and len(y) > 1:
                    p = min(pos, len(y) - 2)
                    y[p], y[p + 1] = y[p + 1], y[p]
            y = "".join(y)
        else:
            y = "".join(rng.choice(alphabet) for _ in range(rng.randint(0, cap)))
        a.append(x)
        b.append(y)
    for la in (0, 1, 2, 31, 32, cap - 1, cap):
        for lb in (0, 1, 2, 31, 32, cap - 1, cap):
            a.append(alphabet[0] * la)
            b.append((alphabet[1] * lb))
            a.append("".join(rng.choice(alphabet) for _ in range(la)))
            b.append("".join(rng.choice(alphabet) for _ in range(lb)))
    ad, ao, _ = _pack(a)
    bd, bo, _ = _pack(b)
    for measure in oracle.MEASURES:
        ints = np.zeros((len(a), 6), dtype=np.int32)
        vals = np.zeros(len(a), dtype=np.float64)
        rc = algos.algos_batch_reg(MEASURE_ID[measure], nbits, len(a), ad.ctypes.data, ao.ctypes.data,
                                   bd.ctypes.data, bo.ctypes.data, ints.ctypes.data, vals.ctypes.data)
        assert rc == 0
        ref, _, ref_ints = oracle.batch(measure, a, b)
        bad = np.nonzero(vals.view(np.uint64) != ref.view(np.uint64))[0]
        assert bad.size == 0, (measure, a[bad[0]], b[bad[0]], vals[bad[0]], ref[bad[0]])
        assert (ints == ref_ints).all(), measure


def test_register_compare_unicode_path(algos, oracle):
    """row_unicode_reg.cuh: tabled string decoded into registers, position masks by compares."""
    from oracle.oracle import _pack, MEASURE_ID

    algos.algos_batch_ureg.restype = ctypes.c_int
    algos.algos_batch_ureg.argtypes = [ctypes.c_int, ctypes.c_int64] + [ctypes.c_void_p] * 6
    rng = random.Random(777)
    a, b = [], []
    while len(a) < 40000:
        x, y = rand_pair(rng, 32)
        if len(x.encode()) <= 32 and len(y.encode()) <= 32:
            a.append(x)
            b.append(y)
    for ch in ("é", "日", "\U0001f600", "a"):
        w = len(ch.encode())
        for la in (0, 1, 2, 32 // w):
            for lb in (0, 1, 32 // w):
                a += [ch * la, ch * la]
                b += [ch * lb, ("z" * lb)]
    ad, ao, _ = _pack(a)
    bd, bo, _ = _pack(b)
    for measure in oracle.MEASURES:
        ints = np.zeros((len(a), 6), dtype=np.int32)
        vals = np.zeros(len(a), dtype=np.float64)
        rc = algos.algos_batch_ureg(MEASURE_ID[measure], len(a), ad.ctypes.data, ao.ctypes.data, bd.ctypes.data,
                                    bo.ctypes.data, ints.ctypes.data, vals.ctypes.data)
        assert rc == 0
        ref, _, ref_ints = oracle.batch(measure, a, b)
        bad = np.nonzero(vals.view(np.uint64) != ref.view(np.uint64))[0]
        assert bad.size == 0, (measure, a[bad[0]], b[bad[0]], vals[bad[0]], ref[bad[0]])
        ib = np.nonzero((ints != ref_ints).any(axis=1))[0]
        assert ib.size == 0, (measure, a[ib[0]], b[ib[0]], ints[ib[0]], ref_ints[ib[0]])


GROUP_MEASURES = {1: ["levenshtein"], 2: ["jaro", "jaro_winkler"], 4: ["jaccard", "sorensen_dice"]}


def check_multi(oracle, call, a, b):
    """call(groups, ints[5][n][6], vals[5][n]) runs a fused row function; every measure of the groups
    must carry the oracle's bits and integer record, the other slots must stay untouched."""
    from oracle.oracle import MEASURE_ID

    n = len(a)
    refs = {m: oracle.batch(m, a, b) for m in oracle.MEASURES}
    for groups in range(1, 8):
        ints = np.full((5, n, 6), -7, dtype=np.int32)
        vals = np.full((5, n), -7.0, dtype=np.float64)
        assert call(groups, ints, vals) == 0
        wanted = [m for g, ms in GROUP_MEASURES.items() if groups & g for m in ms]
        for m in oracle.MEASURES:
            k = MEASURE_ID[m]
            if m not in wanted:
                assert (vals[k] == -7.0).all() and (ints[k] == -7).all(), (groups, m)
                continue
            ref, _, ref_ints = refs[m]
            bad = np.nonzero(vals[k].view(np.uint64) != ref.view(np.uint64))[0]
            assert bad.size == 0, (groups, m, a[bad[0]], b[bad[0]], vals[k][bad[0]], ref[bad[0]])
            ib = np.nonzero((ints[k] != ref_ints).any(axis=1))[0]
            assert ib.size == 0, (groups, m, a[ib[0]], b[ib[0]], ints[k][ib[0]], ref_ints[ib[0]])


@pytest.mark.parametrize("nbits,alphabet", [(5, "abcdefghijklmnopqrstuvwxyz"), (6, "abcxyzABCXYZ_`{"),
                                            (7, "ab yz-'09AZ~\x01\x7f!"), (107, "abcdefghij -'")])
def test_fused_measures_ascii_reg(algos, oracle, nbits, alphabet):
    """row_ascii_reg_multi: one pass over a feeds Myers, Jaro and the multiset together."""
    from oracle.oracle import _pack

    algos.algos_batch_reg_multi.restype = ctypes.c_int
    algos.algos_batch_reg_multi.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64] + [ctypes.c_void_p] * 6
    rng = random.Random(31 + nbits)
    cap = 64 if nbits >= 100 else 32
    a, b = [], []
    for _ in range(6000):
        x = "".join(rng.choice(alphabet) for _ in range(rng.randint(0, cap)))
        if rng.random() < 0.6:
            y = list(x)
            for _ in range(rng.randint(0, 3)):
                op, pos = rng.randint(0, 3), rng.randint(0, len(y))
                if op == 0 and y:
                    y[min(pos, len(y) - 1)] = rng.choice(alphabet)
                elif op == 1 and len(y) < cap:
                    y.insert(pos, rng.choice(alphabet))
                elif op == 2 and y:
                    del y[min(pos, len(y) - 1)]
                elif op == 3 and len(y) > 1:
                    q = min(pos, len(y) - 2)
                    y[q], y[q + 1] = y[q + 1], y[q]
            y = "".join(y)
        else:
            y = "".join(rng.choice(alphabet) for _ in range(rng.randint(0, cap)))
        a.append(x)
        b.append(y)
    for la in (0, 1, 2, 3, 31, 32, cap - 1, cap):
        for lb in (0, 1, 2, 3, 31, 32, cap - 1, cap):
            a.append(alphabet[0] * la)
            b.append(alphabet[1] * lb)
            a.append("".join(rng.choice(alphabet) for _ in range(la)))
            b.append("".join(rng.choice(alphabet) for _ in range(lb)))
    if nbits == 5:
        fx = load_fixture()
        a += [r[1] for r in fx]
        b += [r[2] for r in fx]
    ad, ao, _ = _pack(a)
    bd, bo, _ = _pack(b)
    check_multi(oracle, lambda g, ints, vals: algos.algos_batch_reg_multi(
        g, nbits, len(a), ad.ctypes.data, ao.ctypes.data, bd.ctypes.data, bo.ctypes.data, ints.ctypes.data,
        vals.ctypes.data), a, b)


def test_fused_measures_unicode_reg(algos, oracle):
    """row_unicode_reg_multi: the same fusion on the register-compare path for any script."""
    from oracle.oracle import _pack

    algos.algos_batch_ureg_multi.restype = ctypes.c_int
    algos.algos_batch_ureg_multi.argtypes = [ctypes.c_int, ctypes.c_int64] + [ctypes.c_void_p] * 6
    rng = random.Random(778)
    a, b = [], []
    while len(a) < 12000:
        x, y = rand_pair(rng, 32)
        if len(x.encode()) <= 32 and len(y.encode()) <= 32:
            a.append(x)
            b.append(y)
    for ch in ("é", "日", "\U0001f600", "a"):
        w = len(ch.encode())
        for la in (0, 1, 2, 32 // w):
            for lb in (0, 1, 32 // w):
                a += [ch * la, ch * la]
                b += [ch * lb, ("z" * lb)]
    ad, ao, _ = _pack(a)
    bd, bo, _ = _pack(b)
    check_multi(oracle, lambda g, ints, vals: algos.algos_batch_ureg_multi(
        g, len(a), ad.ctypes.data, ao.ctypes.data, bd.ctypes.data, bo.ctypes.data, ints.ctypes.data,
        vals.ctypes.data), a, b)


@pytest.mark.parametrize("words", [5, 10])
def test_wide_masks_long_ascii_rows(algos, oracle, words):
    """wide_mask.cuh: masks of `words` 32-bit words behind the same step functors -- ASCII strings of up to
    32 * words characters, single measures and every fused group set, against the oracle (bits and integer
    records).  Lengths around every word boundary, empty sides, long common runs, transposed blocks."""
    from oracle.oracle import _pack, MEASURE_ID

    algos.algos_batch_wide.restype = ctypes.c_int
    algos.algos_batch_wide.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64] + [ctypes.c_void_p] * 6
    cap = 32 * words
    alphabet = "ab yz-'09AZ~\x01\x7f!,.#"
    rng = random.Random(4100 + words)
    a, b = [], []
    for _ in range(1500):
        n = rng.choice([rng.randint(0, cap), rng.randint(cap // 2, cap), rng.randint(60, 70)])
        x = "".join(rng.choice(alphabet) for _ in range(n))
        r = rng.random()
        if r < 0.6:
            y = list(x)
            for _ in range(rng.randint(0, 12)):
                op, pos = rng.randint(0, 4), rng.randint(0, len(y))
                if op == 0 and y:
                    y[min(pos, len(y) - 1)] = rng.choice(alphabet)
                elif op == 1 and len(y) < cap:
                    y.insert(pos, rng.choice(alphabet))
                elif op == 2 and y:
                    del y[min(pos, len(y) - 1)]
                elif op == 3 and len(y) > 1:
                    q = min(pos, len(y) - 2)
                    y[q], y[q + 1] = y[q + 1], y[q]
                elif op == 4 and len(y) > 40:  # a block moves: matches far from the diagonal
                    q = rng.randint(0, len(y) - 20)
                    blk = y[q:q + 15]
                    del y[q:q + 15]
                    p2 = rng.randint(0, len(y))
                    y[p2:p2] = blk
            y = "".join(y)
        elif r < 0.8:
            y = "".join(rng.choice(alphabet) for _ in range(rng.randint(0, cap)))
        else:
            y = "".join(rng.choice("ab") for _ in range(rng.randint(0, cap)))
            x = "".join(rng.choice("ab") for _ in range(len(x)))
        a.append(x)
        b.append(y)
    edges = sorted({0, 1, 2, 31, 32, 33, 63, 64, 65, 95, 96, 97, cap // 2 - 1, cap // 2, cap // 2 + 1, cap - 33, cap - 32,
                    cap - 31, cap - 1, cap})
    for la in edges:
        for lb in edges:
            a.append("a" * la)
            b.append("b" * lb)
            a.append("".join(rng.choice(alphabet) for _ in range(la)))
            b.append("".join(rng.choice(alphabet) for _ in range(lb)))
            s = "".join(rng.choice("abc") for _ in range(max(la, lb)))
            a.append(s[:la])
            b.append(s[max(la, lb) - lb:])
    a += ["x" * cap, "xy" * (cap // 2), "x" * cap]
    b += ["x" * cap, "yx" * (cap // 2), "x" * (cap - 1) + "y"]
    ad, ao, _ = _pack(a)
    bd, bo, _ = _pack(b)
    n = len(a)
    for measure in oracle.MEASURES:
        ints = np.zeros((n, 6), dtype=np.int32)
        vals = np.zeros(n, dtype=np.float64)
        rc = algos.algos_batch_wide(words, MEASURE_ID[measure], 0, n, ad.ctypes.data, ao.ctypes.data, bd.ctypes.data,
                                    bo.ctypes.data, ints.ctypes.data, vals.ctypes.data)
        assert rc == 0
        ref, _, ref_ints = oracle.batch(measure, a, b)
        bad = np.nonzero(vals.view(np.uint64) != ref.view(np.uint64))[0]
        assert bad.size == 0, (measure, a[bad[0]], b[bad[0]], vals[bad[0]], ref[bad[0]], ints[bad[0]], ref_ints[bad[0]])
        ib = np.nonzero((ints != ref_ints).any(axis=1))[0]
        assert ib.size == 0, (measure, a[ib[0]], b[ib[0]], ints[ib[0]], ref_ints[ib[0]])
    check_multi(oracle, lambda g, ints, vals: algos.algos_batch_wide(
        words, 0, g, n, ad.ctypes.data, ao.ctypes.data, bd.ctypes.data, bo.ctypes.data, ints.ctypes.data,
        vals.ctypes.data), a, b)
