"""Pins the CPU oracle: reference golden vectors, README table, independent Python twin.

Mirrors the reference's own test strategy (/root/reference/src/expressions/strsim.rs:347-1534:
`compute(a, b)` against a known answer within 1e-8) and adds what those tests do not cover
(non-ASCII, long strings, nulls, threads).  CPU only.
"""
import json
import math
import random
from pathlib import Path

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from conftest import REFERENCE_RS

GOLDEN = Path(__file__).parent / "golden"
THRESHOLD = 1e-8  # strsim.rs:349


def load_fixture():
    rows = []
    for line in (GOLDEN / "reference_vectors.tsv").read_text().splitlines():
        if line.startswith("#") or not line:
            continue
        measure, a, b, exp, hx, ints = line.split("\t")
        rows.append((measure, a, b, float(exp), float.fromhex(hx), [int(x) for x in ints.split(",")]))
    return rows


def test_fixture_counts():
    rows = load_fixture()
    assert len(rows) == 1115
    per = {}
    for r in rows:
        per[r[0]] = per.get(r[0], 0) + 1
    # Lev 76, Jaro 331, JW 526, Jaccard 91, Dice 91
    assert per == {"levenshtein": 76, "jaro": 331, "jaro_winkler": 526, "jaccard": 91, "sorensen_dice": 91}


@pytest.mark.skipif(not REFERENCE_RS.exists(), reason="reference tree absent (GPU box)")
def test_oracle_vs_reference_vectors_parsed_live(oracle):
    """Every `X.test("a", "b", expected)` line of the reference's test module, parsed at test time."""
    import sys

    sys.path.insert(0, str(GOLDEN))
    from make_golden import parse_reference_vectors

    vecs = parse_reference_vectors()
    assert len(vecs) == 1115
    worst = 0.0
    for measure, a, b, exp in vecs:
        v, _ = oracle.pair(measure, a, b)
        worst = max(worst, abs(v - float(exp)))
        assert abs(v - float(exp)) < THRESHOLD, (measure, a, b, v, exp)
    assert worst < 1e-8
    # the committed fixture is exactly this set, in order
    fx = load_fixture()
    assert [(m, a, b) for m, a, b, *_ in fx] == [(m, a, b) for m, a, b, _ in vecs]


def test_oracle_vs_committed_fixture(oracle):
    for measure, a, b, exp, exact, ints in load_fixture():
        v, got = oracle.pair(measure, a, b)
        assert abs(v - exp) < THRESHOLD
        assert v == exact, (measure, a, b, v.hex(), exact.hex())
        assert got.tolist() == ints


def test_python_twin_vs_fixture(oracle):
    """The independent pure-Python derivation agrees BIT-exactly with the C restatement."""
    for measure, a, b, exp, exact, _ in load_fixture():
        assert oracle.PY_TWIN[measure](a, b) == exact, (measure, a, b)


def test_threshold_vectors_pin_f64_order(oracle):
    # strsim.rs:1121 and :1029 -- jaro == 0.7000000000000001 > 0.7, so the Winkler bonus applies
    v, ints = oracle.pair("jaro", "lycon", "laican")
    assert v == (0.6 + 0.5 + 1.0) / 3.0 == 0.7000000000000001
    v, _ = oracle.pair("jaro_winkler", "lycon", "laican")
    assert abs(v - 0.73) < THRESHOLD
    # strsim.rs:1029: also just above the threshold, but the common prefix is empty (l = 0)
    v, ints = oracle.pair("jaro_winkler", "delazouche", "zouche")
    assert v == 0.7000000000000001 and ints[5] == 0


def test_readme_table(oracle):
    table = json.loads((GOLDEN / "readme_table.json").read_text())
    for measure in oracle.MEASURES:
        vals, valid, _ = oracle.batch(measure, table["name_a"], table["name_b"])
        for v, ok, printed, hx in zip(vals, valid, table["printed"][measure], table["oracle_hex"][measure]):
            assert ok == (printed is not None)  # null iff either input null (README.md:69-70)
            if ok:
                assert abs(v - printed) < 5e-7
                assert v == float.fromhex(hx)


def test_spot_values(oracle):
    # SURVEY.md section 9 spot values
    assert oracle.pair("levenshtein", "phillips", "philips")[1].tolist() == [0, 8, 7, 1, 0, 0]
    assert oracle.pair("jaro", "phillips", "philips")[1].tolist() == [0, 8, 7, 7, 0, 0]
    assert oracle.pair("jaro_winkler", "phillips", "philips")[1].tolist() == [0, 8, 7, 7, 0, 4]
    assert oracle.pair("jaccard", "phillips", "philips")[1].tolist() == [0, 8, 7, 7, 8, 0]
    assert oracle.pair("sorensen_dice", "phillips", "philips")[1].tolist() == [0, 8, 7, 7, 15, 0]
    assert oracle.pair("levenshtein", "naïve", "naive")[0] == 0.8
    assert oracle.pair("jaro", "naïve", "naive")[0] == 0.8666666666666667
    assert oracle.pair("jaro_winkler", "naïve", "naive")[0] == 0.8933333333333333
    assert oracle.pair("jaccard", "naïve", "naive")[0] == 4 / 6
    assert oracle.pair("sorensen_dice", "日本語", "日本")[0] == 0.8
    assert oracle.pair("jaro", "日本語", "日本")[0] == 0.8888888888888888
    for m in oracle.MEASURES:  # no normalisation anywhere: precomposed vs decomposed e-acute
        assert oracle.pair(m, "é", "é")[0] == 0.0
    assert oracle.pair("jaccard", "myers", "myres")[0] == 1.0  # strsim.rs:1354 multiset semantics


ALPHABETS = [
    "ab",
    "abcdefghijklmnopqrstuvwxyz",
    "aeiouàéîõüßñ",
    "一二三四五六日本語",
    "aé日\U0001f600\U00010348z",
]


def rand_pair(rng, max_len=40):
    alpha = rng.choice(ALPHABETS)
    a = "".join(rng.choice(alpha) for _ in range(rng.randint(0, max_len)))
    if rng.random() < 0.6:
        b = list(a)
        for _ in range(rng.randint(0, 3)):
            op = rng.randint(0, 3)
            pos = rng.randint(0, len(b)) if b else 0
            if op == 0 and b:
                b[min(pos, len(b) - 1)] = rng.choice(alpha)
            elif op == 1:
                b.insert(pos, rng.choice(alpha))
            elif op == 2 and b:
                del b[min(pos, len(b) - 1)]
            elif op == 3 and len(b) > 1:
                p = min(pos, len(b) - 2)
                b[p], b[p + 1] = b[p + 1], b[p]
        b = "".join(b)
    else:
        b = "".join(rng.choice(alpha) for _ in range(rng.randint(0, max_len)))
    return a, b


def test_c_vs_python_twin_random(oracle):
    rng = random.Random(1234)
    pairs = [rand_pair(rng) for _ in range(3000)]
    pairs += [rand_pair(rng, 150) for _ in range(60)]
    a, b = [p[0] for p in pairs], [p[1] for p in pairs]
    for measure in oracle.MEASURES:
        vals, valid, ints = oracle.batch(measure, a, b)
        assert valid.all()
        for x, y, v, it in zip(a, b, vals, ints):
            assert oracle.PY_TWIN[measure](x, y) == v, (measure, x, y)
            if it[0] == 0:
                assert it[1] == len(x) and it[2] == len(y)


@settings(max_examples=300, deadline=None)
@given(st.text(max_size=30), st.text(max_size=30))
def test_c_vs_python_twin_hypothesis(a, b):
    from oracle import oracle as o

    for measure in o.MEASURES:
        v, _ = o.pair(measure, a, b)
        assert o.PY_TWIN[measure](a, b) == v
        assert 0.0 <= v <= 1.0
    # byte-equality short circuit, empty rules (strsim.rs:128,182-186)
    for measure in o.MEASURES:
        assert o.pair(measure, a, a)[0] == 1.0
        if a:
            assert o.pair(measure, a, "")[0] == 0.0
            assert o.pair(measure, "", a)[0] == 0.0


def test_jaro_integer_floor_of_t(oracle):
    # 59 of the 331 Jaro vectors have odd t: float halving would fail them
    odd = [r for r in load_fixture() if r[0] == "jaro" and r[5][4] % 2 == 1]
    assert len(odd) == 59
    for _, a, b, exp, exact, ints in odd:
        m, t, la, lb = ints[3], ints[4], ints[1], ints[2]
        wrong = (m / la + m / lb + (m - t / 2) / m) / 3.0
        assert abs(wrong - exp) > THRESHOLD


def test_batch_views_threads_and_nulls(oracle):
    pa = pytest.importorskip("pyarrow")
    rng = random.Random(7)
    n = 5000
    a, b = [], []
    for _ in range(n):
        x, y = rand_pair(rng, 30)
        a.append(None if rng.random() < 0.05 else x)
        b.append(None if rng.random() < 0.05 else y)
    A = pa.array(a, type=pa.string_view())
    B = pa.array(b, type=pa.string_view())
    for measure in oracle.MEASURES:
        ref, ref_valid, ref_ints = oracle.batch(measure, a, b)
        for nt in (1, 3, 8):
            out, valid, ints = oracle.batch_views(measure, A, B, n_threads=nt, want_ints=True)
            assert (valid == ref_valid).all()
            assert (out[valid] == ref[valid]).all()
            assert (ints == ref_ints).all()
    # sliced arrays (non-zero offset) and scalar broadcast on either side
    out, valid = oracle.batch_views("jaro_winkler", A.slice(100, 300), B.slice(100, 300))
    ref, ref_valid, _ = oracle.batch("jaro_winkler", a[100:400], b[100:400])
    assert (valid == ref_valid).all() and (out[valid] == ref[valid]).all()
    lit = pa.array(["smith"], type=pa.string_view())
    out, valid = oracle.batch_views("levenshtein", A, lit, n_threads=4)
    ref, ref_valid, _ = oracle.batch("levenshtein", a, ["smith"] * n)
    assert (valid == ref_valid).all() and (out[valid] == ref[valid]).all()
    out, valid = oracle.batch_views("levenshtein", lit, B, n_threads=4)
    assert len(out) == n  # NOT the reference's 1-row quirk (strsim.rs:73)
    with pytest.raises(ValueError):
        oracle.batch_views("jaro", A.slice(0, 10), B.slice(0, 11))
