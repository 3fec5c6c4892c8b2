"""polars_strsim -- drop-in Python surface of the B200-native string-similarity plugin.

Same five functions, same signatures and same registration call as the reference
(/root/reference/polars_strsim/__init__.py:8-69): each builds a lazy `pl.Expr` through
`register_plugin_function(plugin_path=<this directory>, function_name=<name>, args=[expr, other],
is_elementwise=True)`.  Polars dlopen()s the shared object in this directory
(`libpolars_strsim_b200.so`, built from ../csrc for sm_100a) and calls its `_polars_plugin_<name>`
symbols; the arithmetic runs in hand-written CUDA kernels.  There is no CPU fallback.

`polars_strsim.arrow` offers the same five functions over pyarrow arrays (Arrow C Data
Interface), which is what the tests and bench.py drive when Polars is not installed.
"""
from __future__ import annotations

from pathlib import Path

try:  # Polars is the host engine; everything below needs it, `polars_strsim.arrow` does not
    import polars as pl
    from polars.plugins import register_plugin_function

    from polars_strsim.utils import parse_into_expr

    _POLARS_ERROR = None
except ImportError as _e:  # pragma: no cover - this image has no polars
    pl = None
    _POLARS_ERROR = _e

PLUGIN_PATH = Path(__file__).parent


def _plugin(function_name: str, expr, other):
    if pl is None:
        raise ImportError(
            "polars is not installed; use polars_strsim.arrow.%s(a, b) on pyarrow arrays instead"
            % function_name
        ) from _POLARS_ERROR
    expr = parse_into_expr(expr, dtype=pl.Utf8)
    other = parse_into_expr(other, dtype=pl.Utf8)
    return register_plugin_function(
        plugin_path=PLUGIN_PATH,
        function_name=function_name,
        args=[expr, other],
        is_elementwise=True,
    )


def levenshtein(expr, other):
    """1 - edit_distance / max(len) over Unicode scalar values (reference __init__.py:8-16)."""
    return _plugin("levenshtein", expr, other)


def jaro(expr, other):
    """Jaro similarity (reference __init__.py:19-27)."""
    return _plugin("jaro", expr, other)


def jaro_winkler(expr, other):
    """Jaro-Winkler: threshold > 0.7, prefix <= 4, scale 0.1 (reference __init__.py:30-38)."""
    return _plugin("jaro_winkler", expr, other)


def jaccard(expr, other):
    """Jaccard index over character multisets (reference __init__.py:41-49)."""
    return _plugin("jaccard", expr, other)


def sorensen_dice(expr, other):
    """Sorensen-Dice coefficient over character multisets (reference __init__.py:52-60)."""
    return _plugin("sorensen_dice", expr, other)


__all__ = [
    "levenshtein",
    "jaro",
    "jaro_winkler",
    "jaccard",
    "sorensen_dice",
]
