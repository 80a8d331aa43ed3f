"""``pysnptools.util`` pieces on the hot path: ``sub_matrix`` (util/__init__.py:271-393) and ``get_num_threads``."""
import os

import numpy as np

from . import _lib

_CODE = {np.dtype(np.float32): _lib.F32, np.dtype(np.float64): _lib.F64}


def get_num_threads(num_threads=None):
    """bed_reader.get_num_threads semantics (bed.py:34-36): argument, PST_NUM_THREADS, NUM_THREADS, MKL_NUM_THREADS, else all cores.
    The GPU path has no host thread pool; the value is accepted for interface compatibility."""
    if num_threads is not None:
        return num_threads
    for key in ("PST_NUM_THREADS", "NUM_THREADS", "MKL_NUM_THREADS"):
        if key in os.environ:
            return int(os.environ[key])
    return os.cpu_count() or 1


def sub_matrix(val, row_index_list, col_index_list, order="A", dtype=np.float64, num_threads=None):
    """``val[row_index_list][:, col_index_list]`` for 2-D or 3-D arrays, in the given order / dtype, gathered on the GPU."""
    val = np.asarray(val)
    dtype = np.dtype(dtype)
    if order == "A":
        order = "F" if (val.flags["F_CONTIGUOUS"] and not val.flags["C_CONTIGUOUS"]) else "C"
    rows = np.ascontiguousarray(row_index_list, dtype=np.int64).reshape(-1)
    cols = np.ascontiguousarray(col_index_list, dtype=np.int64).reshape(-1)
    shape3 = val.shape if val.ndim == 3 else (val.shape[0], val.shape[1], 1)
    if val.dtype not in _CODE or dtype not in _CODE or not (val.flags["C_CONTIGUOUS"] or val.flags["F_CONTIGUOUS"]):
        raise NotImplementedError("sub_matrix on the GPU handles contiguous float32 / float64 arrays")
    for idx, cnt in ((rows, shape3[0]), (cols, shape3[1])):
        if idx.size and (idx.min() < 0 or idx.max() >= cnt):
            raise IndexError("index out of range for axis of size {0}".format(cnt))
    out_shape = (len(rows), len(cols)) + ((shape3[2],) if val.ndim == 3 else ())
    out = np.empty(out_shape, dtype=dtype, order=order)
    if out.size:
        _lib.require_gpu()
        order_in = _lib.ORDER_C if val.flags["C_CONTIGUOUS"] else _lib.ORDER_F
        _lib.check(_lib.lib.pstb_subset_host(val.ctypes.data, _CODE[val.dtype], order_in, shape3[0], shape3[1], shape3[2], rows.ctypes.data,
                                             len(rows), cols.ctypes.data, len(cols), out.ctypes.data, _CODE[dtype],
                                             _lib.ORDER_C if order == "C" else _lib.ORDER_F))
    return out
