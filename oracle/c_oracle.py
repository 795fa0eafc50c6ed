"""ctypes loader of oracle/libplace_oracle.so (TEST INFRASTRUCTURE ONLY)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libplace_oracle.so")


def build():
    src = os.path.join(HERE, "place_oracle.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        # -march=native is decided on the machine that runs it; rebuild there if the ISA differs
        subprocess.run(["gcc", "-O3", "-fopenmp", "-fPIC", "-shared", "-o", LIB, src, "-lm"], check=True)
    return LIB


def place(codes, S, FP, FN, n_threads=0):
    build()
    L = C.CDLL(LIB)
    L.place_oracle.restype = C.c_int
    L.place_oracle.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_double, C.c_double,
                               C.c_void_p, C.c_int32]
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    S = np.ascontiguousarray(S, dtype=np.uint8)
    soft = np.zeros(S.shape[0])
    rc = L.place_oracle(codes.ctypes.data, codes.shape[0], codes.shape[1], S.ctypes.data, S.shape[0], float(FP), float(FN),
                        soft.ctypes.data, int(n_threads))
    if rc != 0:
        raise MemoryError("place_oracle failed")
    return soft
