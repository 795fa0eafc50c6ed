import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def golden_path(name):
    return os.path.join(GOLDEN, name + ".npz")


def case_inputs(g):
    """(vcf, tree, meta) of a golden case.  Every input file -- the reference's own example_data included,
    copied byte for byte by make_golden.py -- lives under tests/golden/data/, so the same tests run on the
    GPU box, where /root/reference does not exist."""
    import json
    meta = json.loads(str(g["meta"]))
    return os.path.join(REPO, meta["vcf"]), os.path.join(REPO, meta["tree"]), meta


@pytest.fixture(scope="session")
def ctx():
    from scsommerclock_b200 import engine
    return engine.Context(0)
