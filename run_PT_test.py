#!/usr/bin/env python3
"""Drop-in for the reference's command line:

    python run_PT_test.py <VCF_FILE> <NEWICK_TREE_FILE> [-o] [-excl] [-incl] [-w] [-FN] [-FP]

Same arguments and TSV output as cbg-ethz/scSomMerClock's run_PT_test.py; the work
runs on a B200 through scsommerclock_b200."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from scsommerclock_b200.run_PT_test import main  # noqa: E402

if __name__ == "__main__":
    main()
