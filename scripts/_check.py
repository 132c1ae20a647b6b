"""Loader for the CUDA-core cross-check kernels (scripts/csrc/conv_direct.cu -> scripts/libgcc_b200_check.so):
`gcc_conv_direct_bf16` / `gcc_wgrad_direct_bf16` take the argument lists of `gcc_conv_gemm_bf16` /
`gcc_wgrad_gemm_bf16`.  Development probes only; the product library and its public header do not contain them."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcc_b200 import _build, _lib

_so = None


def call(name, *args):
    """Routes *_direct_* names to the check library and everything else to the product library."""
    global _so
    if "_direct_" not in name:
        return _lib.call(name, *args)
    if _so is None:
        _lib.lib()  # binds the CUDA context for this thread
        _so = ctypes.CDLL(_build.build_check())
        protos = _lib.parse_header()
        for direct, gemm in (("gcc_conv_direct_bf16", "gcc_conv_gemm_bf16"), ("gcc_wgrad_direct_bf16", "gcc_wgrad_gemm_bf16")):
            fn = getattr(_so, direct)
            fn.restype, fn.argtypes = protos[gemm][0], protos[gemm][1]
        _so.gcc_check_last_error.restype = ctypes.c_char_p
    rc = getattr(_so, name)(*args)
    if rc != 0:
        raise _lib.GccB200Error("%s failed (%d): %s" % (name, rc, _so.gcc_check_last_error().decode()))
