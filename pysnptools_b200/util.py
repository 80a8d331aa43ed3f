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


def pinned_empty(shape, dtype=np.float64, order="F"):
    """An uninitialised NumPy array in page-locked host memory (``pstb_host_alloc``): as ``out=`` of ``Bed.read`` it is filled
    by direct DMA at PCIe speed.  The memory is released when the array (and every view of it) is garbage collected."""
    import ctypes
    import weakref
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    ptr = _lib.lib.pstb_host_alloc(max(1, count * dtype.itemsize))
    if not ptr:
        raise MemoryError(_lib.last_error())
    raw = (ctypes.c_uint8 * max(1, count * dtype.itemsize)).from_address(ptr)
    weakref.finalize(raw, _lib.lib.pstb_host_free, ptr)          # the ctypes buffer is the base object of every view below
    return np.frombuffer(raw, dtype=dtype, count=count).reshape(shape, order="F" if order in ("F", "A") else "C")


def release_gpu_buffers():
    """Return the buffers the host-buffer entry points keep cached for this thread (chunk ring, SYRK workspace, the device copy
    of the last kernel) to the CUDA driver (``pstb_host_release``)."""
    _lib.check(_lib.lib.pstb_host_release())


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


# ---- intersect_apply (util/__init__.py:18-198): caller-side bookkeeping either side of the hot path ----------------
def intersect_ids(idslist):
    """For id lists [n_k, 2] (None entries allowed) -> int array [n_common, len(idslist)]: row r gives, for each list, the
    position of the r-th id common to all non-None lists (-1 column for a None list)."""
    lookups = []
    for ids in idslist:
        lookups.append(None if ids is None else {tuple(x): k for k, x in enumerate(np.asarray(ids))})
    live = [d for d in lookups if d is not None]
    if not live:
        return np.zeros((0, len(idslist)), dtype=np.int64)
    first = live[0]
    common = [key for key in first if all(key in d for d in live[1:])]           # order of the first non-None list
    out = np.full((len(common), len(idslist)), -1, dtype=np.int64)
    for c, d in enumerate(lookups):
        if d is not None:
            out[:, c] = [d[key] for key in common]
    return out


def _same_ids(iid_list):
    live = [np.asarray(x) for x in iid_list if x is not None]
    return all(a.shape == live[0].shape and np.array_equal(a, live[0]) for a in live[1:])


def intersect_apply(data_list, sort_by_dataset=True, intersect_before_standardize=True, is_test=False):
    """Give every dataset the same individuals in the same order (reference: util/__init__.py:18-173).

    Understood: ``None``; SnpReader-like (``.iid`` + ``[iid_idx, :]``); SnpKernel (the subset is pushed INTO its reader
    before standardizing when ``intersect_before_standardize``); square kernels (``[iid_idx]``); ``{'iid':..,'vals':..}``
    dictionaries (changed in place); ``(val, iid)`` tuples.  Unchanged inputs are returned when the ids already agree.
    """
    from .kernelreader import KernelData, SnpKernel
    if len(data_list) == 0:
        raise Exception("Expect a least one input item")
    iid_list, reindex_list = [], []
    for data in data_list:
        if data is None:
            iid, reindex = None, (lambda d, idx: None)
        elif isinstance(data, SnpKernel):
            iid = data.iid
            if intersect_before_standardize:
                reindex = lambda d, idx: SnpKernel(d.snpreader[idx, :], d.standardizer, block_size=d.block_size) if d._index is None else d[idx]
            else:
                reindex = lambda d, idx: d[idx]
        elif isinstance(data, KernelData):
            iid = data.iid1 if is_test else data.iid0
            reindex = (lambda d, idx: d[:, idx]) if (is_test and not np.array_equal(data.iid0, data.iid1)) else (lambda d, idx: d[idx])
        elif isinstance(data, dict):
            iid = data["iid"]

            def reindex(d, idx):
                d["iid"] = np.asarray(d["iid"])[idx]
                d["vals"] = np.asarray(d["vals"])[idx]
                return d
        elif hasattr(data, "iid"):
            iid, reindex = data.iid, (lambda d, idx: d[idx, :])
        else:
            iid, reindex = data[1], (lambda d, idx: (np.asarray(d[0])[idx], np.asarray(d[1])[idx]))
        iid_list.append(iid)
        reindex_list.append(reindex)
    if _same_ids(iid_list):
        return data_list
    indarr = intersect_ids(iid_list)
    assert indarr.shape[0] > 0, "no individuals remain after intersection, check that ids match in files"
    if sort_by_dataset:
        for c, iid in enumerate(iid_list):
            if iid is not None:
                indarr = indarr[np.argsort(indarr[:, c], kind="stable")]
                break
    return [reindex_list[c](data_list[c], indarr[:, c]) for c in range(len(data_list))]
