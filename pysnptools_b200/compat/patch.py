"""``patch_reference()``: make the UNMODIFIED reference package compute its kinship matrices on the GPU.

With ``pysnptools_b200/compat`` on ``sys.path`` the reference's ``bed_reader`` imports resolve to the CUDA shim, so ``Bed.read``,
``Unit().standardize`` and ``util.sub_matrix`` already run on ``libpst_b200.so``.  The kernel stage, however -- where the reference
spends >99 % of its time for N >~ 10^4 -- is pure Python + NumPy in the reference (the block loop ``SnpReader._read_kernel``,
``pysnptools/snpreader/snpreader.py:623-668``, ending in ``val.dot(val.T)``, ``pysnptools/snpreader/snpdata.py:190-214``), so no
``bed_reader`` symbol can reach it.  This module rebinds those two methods:

* ``SnpReader._read_kernel`` -- for a ``Bed`` (or any nesting of ``reader[iid, sid]`` subsets of one; ``snpreader/_subset.py``)
  with ``Unit`` / ``Beta`` / ``UnitTrained`` / ``BetaTrained`` / ``Identity``: ONE call of ``pstb_snp_kernel_host`` -- the packed
  file bytes stream to the GPU, decode + standardize + tcgen05 SYRK per SNP chunk, K comes back -- and the trained standardizer is
  rebuilt from the per-SNP statistics exactly as ``Unit._merge_trained`` / ``Beta._merge_trained`` would (``unit.py:53-56``,
  ``beta.py:52-63``).  A ``DistributedBed`` (``distributedbed.py:107-207``) is streamed piece by piece, K accumulated on the device.
  Anything else (other readers, other standardizers, ``force_python_only=True``, ``ARRAY_MODULE=cupy``) runs the reference's own code,
  which then lands in the second override for its GEMM.
* ``SnpData._read_kernel`` -- the ``val.dot(val.T)`` of in-memory values: ``pstb_float_kernel``.

Precision follows the dtype the caller asks for, like the reference's own BLAS call does: ``dtype=float32`` -> the tensor-core path
(fp32 accumulation, <= 1e-5 relative Frobenius error, the north_star gate); ``dtype=float64`` (the reference's default) -> with
``float64="exact"`` the float64 FMA path of the library, with ``float64="tensor"`` the tensor-core path converted to float64.

``unpatch_reference()`` restores the originals.  ``launches()`` is ``pstb_launch_count()`` -- tests use it to prove the GPU ran.
"""
import os

import numpy as np

from pysnptools_b200 import _lib

_STATE = {"patched": False, "float64": "tensor", "orig_reader": None, "orig_data": None, "calls": {"fused": 0, "pieces": 0, "float": 0, "reference": 0}}


def launches():
    return int(_lib.lib.pstb_launch_count())


def stats():
    """How often each route was taken since patching: fused (Bed -> pstb_snp_kernel_host), pieces (DistributedBed), float (val.dot(val.T)
    on the GPU), reference (left to the reference's own code)."""
    return dict(_STATE["calls"])


def _ref():
    import pysnptools.snpreader as sr
    import pysnptools.standardizer as st
    from pysnptools.pstreader import PstReader
    from pysnptools.snpreader._subset import _SnpSubset
    return sr, st, PstReader, _SnpSubset


def _resolve_to_root(reader, _SnpSubset, PstReader):
    """(root reader, iid index or None, sid index or None) through any nesting of ``reader[iid, sid]`` (pstreader/_subset.py:55-142)."""
    chain = []
    r = reader
    while isinstance(r, _SnpSubset):
        chain.append(r)
        r = r._internal
    ii = si = None
    for sub in reversed(chain):                                   # from the root outwards
        inner = sub._internal
        a = PstReader._make_sparray_from_sparray_or_slice(inner.row_count, sub._row_indexer).astype(np.int64)
        b = PstReader._make_sparray_from_sparray_or_slice(inner.col_count, sub._col_indexer).astype(np.int64)
        ii = a if ii is None else ii[a]
        si = b if si is None else si[b]
    return r, ii, si


def _spec_of(standardizer, st, sid):
    """(spec, stats or None, kind) for the standardizers the fused kernels implement; None for everything else."""
    t = type(standardizer)
    if t is st.Unit:
        return ("unit",), None, "unit"
    if t is st.Beta:
        return ("beta", float(standardizer.a), float(standardizer.b)), None, "beta"
    if t is st.Identity:
        return ("unit",), np.tile(np.array([[0.0, 1.0]]), (len(sid), 1)), "identity"       # x = dosage; missing -> 0
    if t is st.UnitTrained or t is st.BetaTrained:
        tr_sid = np.asarray(standardizer.sid)
        if len(tr_sid) == len(sid) and np.array_equal(tr_sid, sid):
            stats_ = np.asarray(standardizer.stats, dtype=np.float64)
        elif t is st.UnitTrained:                                   # unittrained.py:55-59: looked up by sid
            if standardizer.sid_to_index is None:
                standardizer.sid_to_index = {s: i for i, s in enumerate(standardizer.sid)}
            stats_ = np.array([standardizer.stats[standardizer.sid_to_index[s]] for s in sid], dtype=np.float64)
        else:                                                       # betatrained.py:53: must match
            raise AssertionError("sid in training and use must be the same and in the same order")
        spec = ("unit",) if t is st.UnitTrained else ("beta", float(standardizer.a), float(standardizer.b))
        return spec, stats_.reshape(len(sid), 2), "trained"
    return None


def _trained(standardizer, kind, st, sid, stats_, dtype):
    if kind == "unit":
        return st.UnitTrained(sid, stats_.astype(dtype))
    if kind == "beta":
        return st.BetaTrained(standardizer.a, standardizer.b, sid, stats_.astype(dtype))
    return standardizer                                             # Identity / trained ones return themselves (identity.py, unittrained.py:67-68)


def _finish(K, order):
    # K is symmetric: its transpose is the F-ordered array with the same values (snpdata.py:201-204 does the same)
    if order == "F":
        return K.T
    return K


def _gpu_route_ok(force_python_only, dtype):
    if force_python_only or os.environ.get("ARRAY_MODULE", "numpy") not in ("", "numpy"):
        return False
    return np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.float64))


def _reader_read_kernel(self, standardizer, block_size=None, order="A", dtype=np.float64, force_python_only=False, view_ok=False,
                        return_trained=False, num_threads=None):
    """Replacement of ``SnpReader._read_kernel`` (snpreader.py:623-668)."""
    sr, st, PstReader, _SnpSubset = _ref()
    dtype = np.dtype(dtype)
    orig = _STATE["orig_reader"]
    if _gpu_route_ok(force_python_only, dtype) and not hasattr(self, "val"):
        root, ii, si = _resolve_to_root(self, _SnpSubset, PstReader)
        is_bed = type(root) is sr.Bed
        is_dbed = type(root).__name__ == "DistributedBed" and hasattr(root, "_merge")
        if (is_bed or is_dbed) and self.iid_count > 0:
            sid = self.sid
            got = _spec_of(standardizer, st, sid)
            if got is not None:
                spec, stats_in, kind = got
                if is_bed:
                    K, stats_ = _bed_kernel(root, ii, si, spec, stats_in, dtype, block_size)
                    _STATE["calls"]["fused"] += 1
                else:
                    K, stats_ = _dbed_kernel(root, ii, si, spec, stats_in, dtype, block_size)
                    _STATE["calls"]["pieces"] += 1
                K = _finish(K, order)
                return (K, _trained(standardizer, kind, st, sid, stats_, dtype)) if return_trained else K
    _STATE["calls"]["reference"] += 1
    return orig(self, standardizer, block_size=block_size, order=order, dtype=dtype, force_python_only=force_python_only, view_ok=view_ok,
                return_trained=return_trained, num_threads=num_threads)


def _chunk_for(block_size, n_iid, n_sid):
    from pysnptools_b200 import device
    chunk = device.default_kernel_chunk(n_iid, n_sid)
    if block_size is not None:
        chunk = min(chunk, max(64, (int(block_size) + 63) // 64 * 64))
    return chunk


def _bed_kernel(bed, ii, si, spec, stats_in, dtype, block_size):
    """``pstb_snp_kernel_host`` on the file behind a reference ``Bed`` (its ``open_bed`` handle is the CUDA shim's)."""
    bed._open_bed_if_needed()
    handle = bed._open_bed
    packed = handle._packed_host()
    n, m = handle.iid_count, handle.sid_count
    ii = None if ii is None else np.ascontiguousarray(ii, dtype=np.int64)
    si = None if si is None else np.ascontiguousarray(si, dtype=np.int64)
    for idx, cnt in ((ii, n), (si, m)):
        if idx is not None and idx.size and (idx.min() < 0 or idx.max() >= cnt):
            raise IndexError("index out of range for axis of size {0}".format(cnt))
    ni, ns = (n if ii is None else len(ii)), (m if si is None else len(si))
    K = np.empty((ni, ni), dtype=dtype)
    stats_ = np.empty((ns, 2), dtype=np.float64)
    use_stats = 0
    if stats_in is not None:
        stats_[...] = stats_in
        use_stats = 1
    mode = _lib.STD_UNIT if spec[0] == "unit" else _lib.STD_BETA
    a, b = (spec[1], spec[2]) if spec[0] == "beta" else (float("nan"), float("nan"))
    _lib.require_gpu()
    exact = dtype == np.float64 and _STATE["float64"] == "exact"
    fn = _lib.lib.pstb_snp_kernel_host_f64 if exact else _lib.lib.pstb_snp_kernel_host
    args = [packed.ctypes.data if m else None, n, m, ii.ctypes.data if ii is not None else None, ni, si.ctypes.data if si is not None else None,
            ns, int(bool(handle.count_A1)), mode, a, b, use_stats, stats_.ctypes.data, K.ctypes.data]
    if exact:
        _lib.check(fn(*args, int(_chunk_for(block_size, ni, ns))))
    else:
        _lib.check(fn(*args, _lib.F64 if dtype == np.float64 else _lib.F32, int(_chunk_for(block_size, ni, ns)), _lib.LOW_TERM_DEFAULT))
    return K, stats_


def _dbed_kernel(dbed, ii, si, spec, stats_in, dtype, block_size):
    """K accumulated over the pieces of a reference ``DistributedBed`` (each piece is a plain ``Bed``; column -> piece routing as in
    ``pstreader/_mergecols.py:105-158``): one packed store on the GPU at a time."""
    import torch
    from pysnptools_b200 import device
    dbed._run_once()
    pieces = list(dbed._merge.reader_list)
    counts = np.array([p.sid_count for p in pieces], dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(counts)])
    n = dbed.iid_count
    sid_all = np.arange(int(starts[-1]), dtype=np.int64) if si is None else np.asarray(si, dtype=np.int64)
    ni = n if ii is None else len(ii)
    piece_of = np.searchsorted(starts, sid_all, side="right") - 1
    low_term = device.low_term_for(len(sid_all), ni, spec)
    K = None
    stats_ = np.empty((len(sid_all), 2), dtype=np.float64)
    for k in np.unique(piece_of):
        where = np.nonzero(piece_of == k)[0]
        piece = pieces[int(k)]
        if hasattr(piece, "local"):                                  # _Distributed1Bed (distributedbed.py:210-268): a Bed once its file is local
            piece._run_once()
            piece = piece.local
        piece._open_bed_if_needed()
        h = piece._open_bed
        store = device.PackedStore.from_host(np.asarray(h._packed_host()), n)
        local = sid_all[where] - starts[k]
        K, st_k = device.snp_kernel(store, ii, local, count_A1=bool(h.count_A1), standardizer=spec,
                                    stats=None if stats_in is None else stats_in[where], chunk=_chunk_for(block_size, ni, len(local)),
                                    K=K, accumulate=K is not None, mirror=False, low_term=low_term)
        stats_[where] = st_k.cpu().numpy()
        del store
    if K is None:
        K = torch.zeros((ni, ni), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib.pstb_mirror_lower(K.data_ptr(), ni, ni, torch.cuda.current_stream().cuda_stream))
    return device.convert_kernel(K, dtype).cpu().numpy(), stats_


def _data_read_kernel(train, standardizer, block_size=None, order="A", dtype=np.float64, force_python_only=False, view_ok=False,
                      return_trained=False, num_threads=None):
    """Replacement of ``SnpData._read_kernel`` (snpdata.py:190-214): ``val.dot(val.T)`` on the GPU."""
    sr, st, PstReader, _SnpSubset = _ref()
    dtype = np.dtype(dtype)
    val = train.val
    if (type(standardizer) is st.Identity and isinstance(val, np.ndarray) and val.dtype == dtype and _gpu_route_ok(force_python_only, dtype)
            and val.ndim == 2 and val.shape[0] > 0 and (val.flags["C_CONTIGUOUS"] or val.flags["F_CONTIGUOUS"])):
        K = _float_kernel_host(val, dtype)
        _STATE["calls"]["float"] += 1
        K = _finish(K, order)
        assert PstReader._array_properties_are_ok(K, order, dtype), "internal error: K is not of the expected order or dtype"
        return (K, standardizer) if return_trained else K
    return _STATE["orig_data"](train, standardizer, block_size=block_size, order=order, dtype=dtype, force_python_only=force_python_only,
                               view_ok=view_ok, return_trained=return_trained, num_threads=num_threads)


def _float_kernel_host(val, dtype):
    import torch
    from pysnptools_b200 import device
    _lib.require_gpu()
    if val.flags["C_CONTIGUOUS"]:
        v = torch.from_numpy(val).cuda()
    else:
        v = torch.from_numpy(val.T).cuda().t()                       # F order: the transposed view is the contiguous one
    if dtype == np.float64 and _STATE["float64"] == "exact":
        return device.float_kernel_f64(v).cpu().numpy()
    K32 = device.float_kernel(v)
    return device.convert_kernel(K32, dtype).cpu().numpy()


def patch_reference(float64="tensor"):
    """Rebind ``SnpReader._read_kernel`` and ``SnpData._read_kernel`` of the importable ``pysnptools`` package to the GPU versions.

    ``float64``: what a ``dtype=float64`` kernel request gets -- ``"tensor"``: the tensor-core path (fp32 accumulation) converted to
    float64; ``"exact"``: the library's float64 path.  ``dtype=float32`` always takes the tensor cores.  Idempotent."""
    if float64 not in ("tensor", "exact"):
        raise ValueError("float64 must be 'tensor' or 'exact'")
    if float64 == "exact" and not hasattr(_lib.lib, "pstb_snp_kernel_host_f64"):
        raise NotImplementedError("this build of libpst_b200.so has no float64 kernel path")
    sr, st, PstReader, _SnpSubset = _ref()
    _STATE["float64"] = float64
    if _STATE["patched"]:
        return
    _STATE["orig_reader"] = sr.SnpReader._read_kernel
    _STATE["orig_data"] = sr.SnpData._read_kernel
    sr.SnpReader._read_kernel = _reader_read_kernel
    sr.SnpData._read_kernel = _data_read_kernel
    _STATE["patched"] = True


def unpatch_reference():
    if not _STATE["patched"]:
        return
    sr, st, PstReader, _SnpSubset = _ref()
    sr.SnpReader._read_kernel = _STATE["orig_reader"]
    sr.SnpData._read_kernel = _STATE["orig_data"]
    _STATE["patched"] = False
