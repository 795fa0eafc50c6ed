"""B200-native hot path of the Poisson Tree (PT) test of cbg-ethz/scSomMerClock.

``run_PT_test`` mirrors the reference's module of the same name (same function
names, CLI and TSV output); the numbers come from hand-written sm_100a CUDA behind
the C ABI in ``include/scsommerclock_b200.h``.
"""
__version__ = "0.1.0"
