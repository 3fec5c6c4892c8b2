"""Expression coercion helper with the behaviour of /root/reference/polars_strsim/utils.py:6-43:
`str` is a column name unless `str_as_lit`, `pl.Expr` passes through, everything else becomes a
literal of the requested dtype."""
from __future__ import annotations

import polars as pl


def parse_into_expr(expr, *, str_as_lit: bool = False, list_as_lit: bool = True, dtype=None) -> "pl.Expr":
    if isinstance(expr, pl.Expr):
        return expr
    if isinstance(expr, str) and not str_as_lit:
        return pl.col(expr)
    if isinstance(expr, list) and not list_as_lit:
        return pl.lit(pl.Series(expr), dtype=dtype)
    return pl.lit(expr, dtype=dtype)
