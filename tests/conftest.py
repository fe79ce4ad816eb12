import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

REFERENCE_DATA = "/root/reference/data"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def have_reference():
    return os.path.isdir(REFERENCE_DATA)


def have_oracle():
    return os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so"))


needs_reference = pytest.mark.skipif(not have_reference(), reason="/root/reference not present on this box")
needs_oracle = pytest.mark.skipif(not have_oracle(), reason="oracle/liboracle.so not built")
