import importlib
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def rtnw():
    """the package (its directory name has hyphens, so it cannot be imported with an import statement)"""
    return importlib.import_module("peter-shirley-ray-tracing-the-next-week_b200")


@pytest.fixture(scope="session")
def ctx(rtnw):
    c = rtnw.Context(0)
    yield c
    c.close()
