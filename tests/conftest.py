"""pytest configuration: the `gpu` marker, import paths, shared helpers."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "polars-strsim_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

REFERENCE_RS = Path("/root/reference/src/expressions/strsim.rs")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o

    o.build()
    return o
