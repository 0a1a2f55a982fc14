"""Build + load the C oracle (oracle/nnmf_ref.c) with gcc.  Test infrastructure only."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "nnmf_ref.c")
_LIB = os.path.join(_HERE, "liboracle_nnmf.so")
_lib = None


def build_oracle(force=False):
    """gcc -O3 -fopenmp oracle/nnmf_ref.c -> oracle/liboracle_nnmf.so (git-ignored, travels with gpurun).

    No -march=native: the .so is built in the CPU container and executed on the GPU box's host."""
    if (not force) and os.path.exists(_LIB) and os.path.getmtime(_LIB) >= os.path.getmtime(_SRC):
        return _LIB
    cmd = ["gcc", "-O3", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-o", _LIB, _SRC, "-lm"]
    subprocess.check_call(cmd)
    return _LIB


def load_oracle():
    global _lib
    if _lib is not None:
        return _lib
    path = build_oracle()
    lib = ctypes.CDLL(path)
    d = ctypes.POINTER(ctypes.c_double)
    ip = ctypes.POINTER(ctypes.c_int)
    lib.nnmf_ref.restype = ctypes.c_int
    lib.nnmf_ref.argtypes = [d, ctypes.c_int, ctypes.c_int, ctypes.c_int, d, d, d, d, ctypes.c_int,
                             ctypes.c_double, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                             ctypes.c_int, d, d, d, d, ip, ip]
    lib.nnlm_ref.restype = ctypes.c_long
    lib.nnlm_ref.argtypes = [d, d, d, ctypes.c_int, ctypes.c_int, ctypes.c_int, d, ctypes.c_int,
                             ctypes.c_double, ctypes.c_int, ctypes.c_int]
    lib.oracle_num_threads.restype = ctypes.c_int
    lib.oracle_set_variant.argtypes = [ctypes.c_int]
    lib.oracle_set_variant.restype = None
    _lib = lib
    return lib
