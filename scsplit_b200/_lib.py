"""ctypes binding of libscsplit_b200.so (include/scsplit_b200.h).

The library is the product: if it is missing or cannot be loaded this module raises -- there is no
CPU implementation to fall back to.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int8, c_int32, c_int64, c_uint8, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# SCSPLIT_B200_LIB points the loader at another build of the same ABI (kernel experiments); the default is the in-tree library
LIB_PATH = os.environ.get("SCSPLIT_B200_LIB") or os.path.join(HERE, "libscsplit_b200.so")

SCS_P, SCS_W, SCS_L, SCS_LPS, SCS_LLHIST, SCS_STATS = range(6)
SCS_MAX_K = 32
SCS_MAX_ITER = 300

# name -> (restype, argtypes); mirrors include/scsplit_b200.h one to one
SIGNATURES = {
    "scs_abi_version": (c_int, []),
    "scs_last_error": (c_char_p, []),
    "scs_device_count": (c_int, [POINTER(c_int)]),
    "scs_create": (c_int, [POINTER(c_void_p), c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                           c_void_p, c_int]),
    "scs_destroy": (c_int, [c_void_p]),
    "scs_layout_info": (c_int, [c_void_p, POINTER(c_int64)]),
    "scs_kalt": (c_int, [c_void_p, c_void_p]),
    "scs_seed_af": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "scs_pattern_counts": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "scs_em_footprint": (c_int, [c_void_p, c_int, POINTER(c_int64)]),
    "scs_em_begin": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int]),
    "scs_em_iterate": (c_int, [c_void_p, c_int, c_int]),
    "scs_em_capture": (c_int, [c_void_p, c_int]),
    "scs_em_last_ms": (c_int, [c_void_p, POINTER(c_double)]),
    "scs_em_run": (c_int, [c_void_p]),
    "scs_estep": (c_int, [c_void_p]),
    "scs_mstep": (c_int, [c_void_p]),
    "scs_em_status": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "scs_em_info": (c_int, [c_void_p, POINTER(c_int64)]),
    "scs_em_get": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "scs_em_set": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "scs_em_end": (c_int, [c_void_p]),
    "scs_doublet_scores": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "scs_refine_scores": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "scs_cluster_counts": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "scs_comm_unique_id": (c_int, [c_void_p]),
    "scs_comm_init": (c_int, [c_void_p, c_void_p, c_int, c_int]),
    "scs_csv_parse": (c_int, [c_char_p, c_int, POINTER(c_void_p)]),
    "scs_csv_shape": (c_int, [c_void_p, POINTER(c_int64)]),
    "scs_csv_fetch": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "scs_csv_free": (c_int, [c_void_p]),
    "scs_csv_write_f64": (c_int, [c_char_p, c_char_p, c_char_p, c_int64, c_int64, c_void_p, c_int]),
    "scs_launch_count": (c_int, [c_void_p, POINTER(c_int64)]),
    "scs_set_timing": (c_int, [c_void_p, c_int]),
    "scs_kernel_times": (c_int, [c_void_p, POINTER(c_double), c_int]),
    "scs_set_variant": (c_int, [c_void_p, c_int]),
    "scs_selftest_exp2": (c_int, [c_void_p, c_void_p, c_void_p, c_int64]),
}

_lib = None


class ScsError(RuntimeError):
    pass


def load():
    """Load the shared library (once). Raises if it was not built: run `python -m scsplit_b200.build`."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ScsError("%s not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                       "scsplit_b200 has no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise ScsError(load().scs_last_error().decode("utf-8", "replace"))


def device_count():
    n = c_int(0)
    rc = load().scs_device_count(ctypes.byref(n))
    return n.value if rc == 0 else 0
