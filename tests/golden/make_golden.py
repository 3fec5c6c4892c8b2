"""Regenerates tests/golden/reference_vectors.tsv and tests/golden/readme_table.json.

Run HERE (the container that has /root/reference); the GPU box has no reference tree, so the
fixtures travel instead.  For every known-answer vector in the reference's unit tests
(/root/reference/src/expressions/strsim.rs:371-1534) this records the inputs, the reference's
printed expectation (8 significant digits, tolerance 1e-8, strsim.rs:349) and the oracle's exact
result (f64 as hex + integer intermediates), after asserting the oracle meets the expectation.

    python tests/golden/make_golden.py
"""
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as o  # noqa: E402

RS = Path("/root/reference/src/expressions/strsim.rs")
PREFIX = {"lev": "levenshtein", "j": "jaro", "jw": "jaro_winkler", "jac": "jaccard", "sd": "sorensen_dice"}
PAT = re.compile(r'^\s*(lev|jw|jac|sd|j)\.test\("([^"]*)", "([^"]*)", ([0-9.]+)\);')


def parse_reference_vectors(path=RS):
    out = []
    for line in path.read_text().splitlines():
        m = PAT.match(line)
        if m:
            out.append((PREFIX[m.group(1)], m.group(2), m.group(3), m.group(4)))
    return out


def main():
    vecs = parse_reference_vectors()
    assert len(vecs) == 1115, len(vecs)
    lines = ["# measure\ta\tb\treference_expected\toracle_f64_hex\tflag,la,lb,x0,x1,x2"]
    for measure, a, b, exp in vecs:
        v, ints = o.pair(measure, a, b)
        assert abs(v - float(exp)) < 1e-8, (measure, a, b, v, exp)
        lines.append("\t".join([measure, a, b, exp, float(v).hex(), ",".join(map(str, ints.tolist()))]))
    (Path(__file__).parent / "reference_vectors.tsv").write_text("\n".join(lines) + "\n")

    # README.md:41-70 (== demo.py:4-8): the 6-row frame and its printed table
    table = {
        "name_a": ["phillips", "phillips", "", "", None, None],
        "name_b": ["phillips", "philips", "phillips", "", "phillips", None],
        "printed": {  # 6 significant digits as polars prints them, README.md:65-70
            "levenshtein": [1.0, 0.875, 0.0, 1.0, None, None],
            "jaro": [1.0, 0.958333, 0.0, 1.0, None, None],
            "jaro_winkler": [1.0, 0.975, 0.0, 1.0, None, None],
            "jaccard": [1.0, 0.875, 0.0, 1.0, None, None],
            "sorensen_dice": [1.0, 0.933333, 0.0, 1.0, None, None],
        },
        "oracle_hex": {},
    }
    for measure in o.MEASURES:
        vals, valid, _ = o.batch(measure, table["name_a"], table["name_b"])
        table["oracle_hex"][measure] = [float(v).hex() if ok else None for v, ok in zip(vals, valid)]
        for v, ok, p in zip(vals, valid, table["printed"][measure]):
            assert (p is None) == (not ok)
            assert p is None or abs(v - p) < 5e-7
    (Path(__file__).parent / "readme_table.json").write_text(json.dumps(table, indent=1) + "\n")
    print("wrote", len(vecs), "vectors")


if __name__ == "__main__":
    main()
