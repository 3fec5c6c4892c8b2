import sys; sys.path[:0]=['.','polars-strsim_b200','tests']
import pyarrow as pa, numpy as np
from polars_strsim import _native
from test_oracle import load_fixture
rows=[r for r in load_fixture() if r[0]=="levenshtein"]
a=[r[1] for r in rows]; b=[r[2] for r in rows]
print(a[:8], b[:8], max(len(x) for x in a+b))
v, valid, n, ints = _native.compute_host("levenshtein", pa.array(a, type=pa.string_view()), pa.array(b, type=pa.string_view()), debug=True)
print(v[:10], ints[:10])
