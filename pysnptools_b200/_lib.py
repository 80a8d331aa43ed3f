"""ctypes binding of ``libpst_b200.so`` (C ABI declared in ``include/pst_b200.h``).

The library is the product: there is no CPU fallback.  If the shared object is missing the import
fails loudly with the build command.
"""
import ctypes
import os
from ctypes import Structure, c_char_p, c_double, c_float, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PSTB_LIB_PATH") or os.path.join(_HERE, "libpst_b200.so")   # override: kernel experiments only

F32, F64, I8 = 0, 1, 2
ORDER_F, ORDER_C = 0, 1
STD_NONE, STD_UNIT, STD_BETA = 0, 1, 2
LOW_TERM_DEFAULT, LOW_TERM_FP16, LOW_TERM_FP8, LOW_TERM_AUTO = -1, 0, 1, 2


class Axis(Structure):
    """``pstb_axis``: device uint32 index vector, or start + k*step when ``idx`` is NULL."""
    _fields_ = [("idx", c_void_p), ("start", c_int64), ("step", c_int64), ("n", c_int64)]


class PstB200Error(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "pysnptools_b200: {0} not found. Build it with `bash pysnptools_b200/csrc/build.sh` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.".format(LIB_PATH))
    lib = ctypes.CDLL(LIB_PATH)
    sig = {
        "pstb_version": (c_int, []),
        "pstb_last_error": (c_char_p, []),
        "pstb_sm_count": (c_int, []),
        "pstb_packed_ld": (c_int64, [c_int64]),
        "pstb_launch_count": (c_int64, []),
        "pstb_host_alloc": (c_void_p, [c_int64]),
        "pstb_host_free": (c_int, [c_void_p]),
        "pstb_host_release": (c_int, []),
        "pstb_numa_bind": (c_int, [c_int]),
        "pstb_decode": (c_int, [c_void_p, c_int64, c_int64, c_int64, Axis, Axis, c_int, c_void_p, c_int, c_int, c_void_p]),
        "pstb_decode_standardize": (c_int, [c_void_p, c_int64, c_int64, c_int64, Axis, Axis, c_int, c_int, c_double, c_double,
                                            c_int, c_void_p, c_void_p, c_int, c_int, c_void_p]),
        "pstb_standardize_work_bytes": (c_int64, [c_int64]),
        "pstb_standardize": (c_int, [c_void_p, c_int, c_int, c_int64, c_int64, c_int, c_double, c_double, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p]),
        "pstb_subset": (c_int, [c_void_p, c_int, c_int, c_int64, c_int64, c_int64, Axis, Axis, c_void_p, c_int, c_int, c_void_p]),
        "pstb_pack": (c_int, [c_void_p, c_int, c_int, c_int64, c_int64, c_int, c_void_p, c_int64, c_void_p, c_void_p]),
        "pstb_kernel_workspace_bytes": (c_int64, [c_int64, c_int64]),
        "pstb_snp_kernel": (c_int, [c_void_p, c_int64, c_int64, c_int64, Axis, Axis, c_int, c_int, c_double, c_double, c_int,
                                    c_void_p, c_void_p, c_int, c_int, c_void_p, c_int64, c_int64, c_int, c_void_p]),
        "pstb_kernel_tile_count": (c_int64, [c_int64, c_int, c_int]),
        "pstb_kernel_tile_coords": (c_int, [c_int64, c_int, c_int, c_void_p]),
        "pstb_snp_kernel_tiles": (c_int, [c_void_p, c_int64, c_int64, c_int64, Axis, Axis, c_int, c_int, c_double, c_double, c_int,
                                          c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int64, c_int64, c_int, c_void_p]),
        "pstb_snp_kernel_tiles_band": (c_int, [c_void_p, c_int64, c_int64, c_int64, Axis, Axis, c_int, c_int, c_double, c_double, c_int,
                                               c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int64, c_int64, c_int, c_int64, c_int64, c_int, c_int,
                                               c_void_p]),
        "pstb_kernel_from_tiles_range": (c_int, [c_void_p, c_int64, c_int, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
        "pstb_kernel_workspace_rank1": (c_void_p, [c_void_p, c_int64, c_int64]),
        "pstb_resolve_low_term": (c_int, [c_int, c_int64, c_int64, c_int]),
        "pstb_set_syrk_low_term": (c_int, [c_int]),
        "pstb_kernel_from_tiles": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
        "pstb_cross_kernel_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64]),
        "pstb_snp_cross_kernel": (c_int, [c_void_p, c_int64, c_int64, c_int64, Axis, Axis, c_int,
                                          c_void_p, c_int64, c_int64, c_int64, Axis, Axis, c_int,
                                          c_int, c_double, c_double, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int64, c_int, c_void_p]),
        "pstb_float_kernel": (c_int, [c_void_p, c_int, c_int, c_int64, c_int64, c_void_p, c_int, c_int, c_void_p, c_int64, c_int64, c_void_p]),
        "pstb_kernel_f64_workspace_bytes": (c_int64, [c_int64, c_int64]),
        "pstb_snp_kernel_f64": (c_int, [c_void_p, c_int64, c_int64, c_int64, Axis, Axis, c_int, c_int, c_double, c_double, c_int,
                                        c_void_p, c_void_p, c_int, c_int, c_void_p, c_int64, c_int64, c_void_p]),
        "pstb_float_kernel_f64": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_int, c_int, c_void_p]),
        "pstb_mirror_lower_f64": (c_int, [c_void_p, c_int64, c_int64, c_void_p]),
        "pstb_snp_kernel_host_f64": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_double,
                                             c_double, c_int, c_void_p, c_void_p, c_int64]),
        "pstb_syrk_planes": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int, c_float, c_void_p]),
        "pstb_mirror_lower": (c_int, [c_void_p, c_int64, c_int64, c_void_p]),
        "pstb_convert_kernel": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_double, c_void_p]),
        "pstb_read_host": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_double,
                                   c_double, c_int, c_void_p, c_void_p, c_int, c_int]),
        "pstb_snp_kernel_host": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_double,
                                         c_double, c_int, c_void_p, c_void_p, c_int, c_int64, c_int]),
        "pstb_standardize_host": (c_int, [c_void_p, c_int, c_int, c_int64, c_int64, c_int, c_double, c_double, c_int, c_int, c_void_p]),
        "pstb_subset_host": (c_int, [c_void_p, c_int, c_int, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                     c_void_p, c_int, c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    return lib, sorted(sig)


lib, EXPORTS = _load()


def last_error():
    return (lib.pstb_last_error() or b"").decode("utf-8", "replace")


def check(rc):
    if rc != 0:
        msg = last_error()
        if "out of range" in msg or "outside" in msg:
            raise IndexError(msg)
        raise PstB200Error(msg)


def require_gpu():
    if lib.pstb_sm_count() <= 0:
        raise PstB200Error("pysnptools_b200 needs a CUDA device (B200, sm_100a); none is visible and there is no CPU fallback")
