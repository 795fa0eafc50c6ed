"""ctypes binding of the C-ABI library (include/scsommerclock_b200.h).

This is the reference-side stub a maintainer would add: it
loads ``libscs_b200.so`` from the package directory and fails loudly if the
library is missing -- there is no CPU fallback behind it.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libscs_b200.so")

SCS_OK = 0
GT_WT, GT_HET, GT_HOM, GT_MISSING = 0, 1, 2, 3


class ScsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"scsommerclock_b200 error {code}: {msg}")
        self.code = code


_lib = None

_SIGNATURES = {
    # name: (restype, argtypes)
    "scs_version": (C.c_char_p, []),
    "scs_last_error": (C.c_char_p, []),
    "scs_ctx_create": (C.c_int, [C.c_int32, C.POINTER(C.c_void_p)]),
    "scs_ctx_destroy": (C.c_int, [C.c_void_p]),
    "scs_ctx_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_int32]),
    "scs_upload_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "scs_placement_create": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "scs_placement_create_batch": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "scs_placement_run_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scs_placement_destroy": (C.c_int, [C.c_void_p]),
    "scs_placement_set_genotypes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "scs_placement_set_genotypes_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "scs_clade_intervals": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scs_placement_bind_genotypes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "scs_placement_prepare": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "scs_placement_prepare_result": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.c_void_p]),
    "scs_placement_set_clades": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "scs_placement_set_trees": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scs_placement_set_clades_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "scs_placement_run": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]),
    "scs_placement_run_host": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]),
    "scs_placement_device_result": (C.c_void_p, [C.c_void_p]),
    "scs_placement_last_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "scs_placement_prepare_time": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "scs_placement_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_int32]),
    "scs_placement_debug_counts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "scs_placement_debug_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "scs_map_mutations": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32,
                                    C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_double, C.c_void_p]),
    "scs_place_percell": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                    C.c_double, C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.c_void_p]),
    "scs_vcf_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]),
    "scs_vcf_destroy": (C.c_int, [C.c_void_p]),
    "scs_vcf_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_int32]),
    "scs_vcf_device_genotypes": (C.c_void_p, [C.c_void_p]),
    "scs_vcf_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "scs_gt_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int32,
                                C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.c_void_p]),
    "scs_branch_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_double,
                                     C.c_double, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    "scs_lrt_poisson_batch": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
    "scs_lrt_plan_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "scs_lrt_plan_destroy": (C.c_int, [C.c_void_p]),
    "scs_lrt_plan_set_weights": (C.c_int, [C.c_void_p, C.c_void_p]),
    "scs_lrt_plan_set_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "scs_lrt_plan_run": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scs_lrt_plan_run_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scs_lrt_plan_device_result": (C.c_void_p, [C.c_void_p]),
    "scs_lrt_plan_create_uniform": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "scs_lrt_plan_device_inputs": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "scs_tree_parse_batch": (C.c_int, [C.c_int32, C.c_char_p, C.c_void_p, C.c_int32, C.c_int32, C.c_char_p, C.c_void_p, C.c_char_p, C.c_int32,
                                       C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "scs_gt_reduce_groups": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "scs_sweep_model": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "scs_sweep_model_workspace": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "scs_format_pt_rows": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """The loaded library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ScsError(-4, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(nvcc, sm_100a). The PT hot path has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != SCS_OK:
        raise ScsError(rc, lib().scs_last_error().decode("utf-8", "replace"))


def ptr(a):
    """Host pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)
