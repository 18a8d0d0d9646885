import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the reference tree (/root/reference, or its unmodified install under baseline/_ref)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isfile("/root/reference/protonoc/simulator/model.py") or \
        os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "protonoc", "simulator", "model.py"))
    skip_ref = pytest.mark.skip(reason="reference tree not present")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)
