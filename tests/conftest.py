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
    """(vcf, tree) paths of a golden case.  The reference's example_data is not in
    this repo (it lives in /root/reference, absent on the GPU box); its inputs are
    replaced there by the fixture's recorded genotype matrix."""
    import json
    meta = json.loads(str(g["meta"]))
    vcf, tree = meta["vcf"], meta["tree"]
    if vcf.startswith("reference/"):
        vcf_p, tree_p = os.path.join("/root", vcf), os.path.join("/root", tree)
        if not os.path.exists(tree_p) and "tree_text" in g.files:
            d = os.path.join(GOLDEN, "_tmp", meta["case"])
            os.makedirs(d, exist_ok=True)
            tree_p = os.path.join(d, str(g["tree_basename"]))
            with open(tree_p, "w") as f:
                f.write(str(g["tree_text"]))
        return vcf_p, tree_p, meta
    return os.path.join(REPO, vcf), os.path.join(REPO, tree), meta


@pytest.fixture(scope="session")
def ctx():
    from scsommerclock_b200 import engine
    return engine.Context(0)
