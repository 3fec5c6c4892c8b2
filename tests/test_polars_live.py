"""The drop-in under a LIVE Polars, when one is installed: the reference's own demo
(/root/reference/demo.py:1-18) through `polars_strsim`'s five functions -- `register_plugin_function` makes
Polars dlopen() libpolars_strsim_b200.so and call its `_polars_plugin_*` symbols -- against the printed table
of the reference's README (README.md:58-72, committed as tests/golden/readme_table.json).

This image ships no `polars` (and the GPU boxes use the same image), so the test SKIPS here; it is the
check to run first wherever Polars >= 1.0 and a B200 are both at hand.  Until then the plugin symbols are
exercised through fabricated `SeriesExport`s (bench_support/plugin_driver.py, tests/test_abi.py,
tests/test_gpu_parity.py)."""
import json
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

pl = pytest.importorskip("polars", reason="no Polars in this image: the plugin symbols are driven through ctypes instead")


@pytest.mark.gpu
def test_reference_demo_under_live_polars():
    import polars_strsim
    from polars_strsim import jaccard, jaro, jaro_winkler, levenshtein, sorensen_dice

    table = json.loads((ROOT / "tests" / "golden" / "readme_table.json").read_text())
    df = pl.DataFrame({"name_a": table["name_a"], "name_b": table["name_b"]})
    out = df.with_columns(
        levenshtein=levenshtein("name_a", "name_b"),
        jaro=jaro("name_a", "name_b"),
        jaro_winkler=jaro_winkler("name_a", "name_b"),
        jaccard=jaccard("name_a", "name_b"),
        sorensen_dice=sorensen_dice("name_a", "name_b"),
    )
    assert polars_strsim.__all__ == ["levenshtein", "jaro", "jaro_winkler", "jaccard", "sorensen_dice"]
    for measure, expected in table["printed"].items():
        got = out[measure].to_list()
        assert out[measure].dtype == pl.Float64
        for g, e in zip(got, expected):
            assert (g is None) == (e is None), (measure, got, expected)
            if e is not None:
                assert abs(g - e) < 1e-6, (measure, got, expected)  # the README prints six decimals
    # a literal on the right, the lazy engine, and a shape error that must surface as a Polars error
    lit = df.lazy().select(jaro_winkler("name_a", pl.lit("phillips"))).collect()
    assert lit.height == df.height
    with pytest.raises(Exception, match="same length"):
        pl.select(jaro(pl.Series(["a", "b", "c"]), pl.Series(["a", "b"])))
