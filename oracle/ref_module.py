"""The UNMODIFIED reference module as a CPU arm (TEST / BENCH INFRASTRUCTURE ONLY -- the product never imports this).

The reference's PT test is one Python file (run_PT_test.py).  It cannot be pip-installed (no setup.py / pyproject) and
/root/reference does not exist on the GPU box, so the recipe is a byte-for-byte copy:

    make()  copies /root/reference/run_PT_test.py -> oracle/_ref/run_PT_test.py   (build container only; oracle/_ref/ is
            git-ignored -- no reference source enters the history -- but travels to the GPU box with the snapshot)
    load()  imports oracle/_ref/run_PT_test.py exactly like tests/golden/make_golden.py imports the original:
              * `from ete3 import Tree` is served by scsommerclock_b200.tree.Tree (ete3 is not installable here; the
                stand-in restates the ete3 3.1.x TreeNode methods the reference calls),
              * scipy.optimize.minimize is wrapped in np.errstate(under='ignore') (the reference's global
                np.seterr(all='raise') makes SciPy 1.18's trust-constr raise on harmless internal underflow).
            Nothing else is touched: map_mutations_gt (:386-448), get_LRT_poisson (:638-700) ... are the reference's own.

`python -m oracle.ref_module` runs make().  __graft_entry__.build() calls make() when /root/reference is present.
"""
import filecmp
import importlib.util
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
SRC = "/root/reference/run_PT_test.py"
DST_DIR = os.path.join(HERE, "_ref")
DST = os.path.join(DST_DIR, "run_PT_test.py")


def make():
    """Copy the reference's one source file for this path into oracle/_ref/.  Returns the path, or None when
    /root/reference is absent (the GPU box: the prebuilt copy is used as it is)."""
    if not os.path.exists(SRC):
        return DST if os.path.exists(DST) else None
    os.makedirs(DST_DIR, exist_ok=True)
    if not os.path.exists(DST) or not filecmp.cmp(SRC, DST, shallow=False):
        shutil.copyfile(SRC, DST)
    return DST


def available():
    return os.path.exists(DST)


def load():
    """Import oracle/_ref/run_PT_test.py (see the module docstring for the two environment shims)."""
    if not available():
        raise FileNotFoundError("oracle/_ref/run_PT_test.py is missing: run `python -m oracle.ref_module` where /root/reference exists")
    if REPO not in sys.path:
        sys.path.insert(0, REPO)
    import numpy as np
    from scsommerclock_b200 import tree as T
    ete3 = types.ModuleType("ete3")
    ete3.Tree = T.Tree
    parser = types.ModuleType("ete3.parser")
    newick = types.ModuleType("ete3.parser.newick")
    newick.NewickError = T.NewickError
    parser.newick = newick
    ete3.parser = parser
    sys.modules.update({"ete3": ete3, "ete3.parser": parser, "ete3.parser.newick": newick})
    spec = importlib.util.spec_from_file_location("ref_run_PT_test", DST)
    ref = importlib.util.module_from_spec(spec)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)      # the reference's regex literals predate Python 3.12's escape warnings
        spec.loader.exec_module(ref)
    _min = ref.minimize

    def minimize_no_underflow_trap(*a, **k):
        with np.errstate(under="ignore"):
            return _min(*a, **k)

    ref.minimize = minimize_no_underflow_trap
    return ref


if __name__ == "__main__":
    print(make())
