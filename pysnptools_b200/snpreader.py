"""SnpReader / Bed / SnpData: the reference's ``pysnptools.snpreader`` surface for the hot path, backed by CUDA.

Mirrors ``Bed(...)`` (snpreader/bed.py:75-108), ``Bed._read`` (bed.py:318-345), ``SnpReader.read`` (snpreader.py:419-478),
``reader[iid_idx, sid_idx]`` (snpreader.py:522-526; index composition pstreader/_subset.py:55-142),
``read_kernel`` (snpreader.py:528-561, 623-668), ``SnpData`` (snpdata.py:67-86, 138-214) and ``Bed.write`` (bed.py:229-316).
Where the reference calls the Rust ``bed_reader`` this module calls ``libpst_b200.so``; there is no CPU path.
"""
import os
import warnings

import numpy as np

from . import _lib
from .standardizer import Identity, Standardizer, Unit, _is_tensor, _no_python_path

plink_chrom_map = {"X": 23, "Y": 24, "XY": 25, "MT": 26}
_DT_CODE = {np.dtype(np.float32): _lib.F32, np.dtype(np.float64): _lib.F64, np.dtype(np.int8): _lib.I8}


# ---- index resolution (pstreader/pstreader.py:617-654, _subset.py:114-142) --------------------------------
def _resolve_indexer(indexer, count):
    """slice | int | bool mask | int sequence -> None (everything, in order) or a non-negative int64 vector."""
    if indexer is None:
        return None
    if isinstance(indexer, slice):
        if indexer == slice(None):
            return None
        return np.arange(count, dtype=np.int64)[indexer]
    if isinstance(indexer, (int, np.integer)):
        indexer = [int(indexer)]
    arr = np.asarray(indexer)
    if arr.size == 0:
        return np.zeros(0, dtype=np.int64)
    if arr.dtype == bool:
        if arr.shape != (count,):
            raise IndexError("boolean index has length {0}, axis has {1}".format(arr.size, count))
        return np.nonzero(arr)[0].astype(np.int64)
    arr = arr.astype(np.int64).reshape(-1)
    arr = np.where(arr < 0, arr + count, arr)
    if arr.min() < 0 or arr.max() >= count:
        raise IndexError("index out of range for axis of size {0}".format(count))
    return arr


def _compose(outer, inner):
    """Index vector of ``reader[outer][inner]`` in the root reader's coordinates (None = all)."""
    if inner is None:
        return outer
    if outer is None:
        return inner
    return outer[inner]


def _symmetric_to_host(K, order):
    """Device kernel -> NumPy in the requested order.  K is exactly symmetric (the upper triangle is a copy of the lower
    one), so its transpose view IS the F-ordered matrix: no 8*N^2-byte re-layout on the host (snpdata.py:207-209)."""
    host = K.cpu().numpy()
    return host.T if order == "F" else host


def _order_code(order):
    if order in ("F", "A"):
        return _lib.ORDER_F
    if order == "C":
        return _lib.ORDER_C
    raise ValueError("order must be 'F', 'C' or 'A'")


class SnpReader(object):
    """Base of Bed / SnpData / subsets: labels + lazy subsetting + read + read_kernel."""

    # --- labels ---
    @property
    def iid(self):
        return self.row

    @property
    def sid(self):
        return self.col

    @property
    def pos(self):
        return self.col_property

    @property
    def iid_count(self):
        return len(self.row)

    @property
    def sid_count(self):
        return len(self.col)

    @property
    def row_count(self):
        return self.iid_count

    @property
    def col_count(self):
        return self.sid_count

    @property
    def shape(self):
        return (self.iid_count, self.sid_count)

    @property
    def row_property(self):
        """SnpReaders carry no per-iid properties: an empty [iid_count, 0] array (snpreader.py:393-397)."""
        return np.empty((self.iid_count, 0))

    @property
    def val_shape(self):
        return None

    def iid_to_index(self, list):
        """Indices of the given iids; KeyError for an unknown one (pstreader.py:339-352)."""
        lookup = {tuple(x): i for i, x in enumerate(self.iid)}
        return np.array([lookup[tuple(x)] for x in list], dtype=np.int64)

    def sid_to_index(self, list):
        lookup = {x: i for i, x in enumerate(self.sid)}
        return np.array([lookup[x] for x in list], dtype=np.int64)

    row_to_index = iid_to_index
    col_to_index = sid_to_index

    def copyinputs(self, copier):
        """Cluster runners' file-staging hook (snpreader.py:563-565): readers with files override it."""
        pass

    # --- subsetting ---
    def __getitem__(self, iid_indexer_and_snp_indexer):
        iid_indexer, sid_indexer = iid_indexer_and_snp_indexer
        return _SnpSubset(self, iid_indexer, sid_indexer)

    # --- reading ---
    def read(self, order="F", dtype=np.float64, force_python_only=False, view_ok=False, num_threads=None,
             _require_float32_64=True, to_device=False, standardizer=None, return_trained=False, out=None):
        """Read into a :class:`SnpData`.

        Extensions over the reference signature: ``to_device=True`` keeps ``val`` as a CUDA tensor; ``standardizer=Unit()``
        (or Beta / a trained one) fuses ``read(...).standardize(standardizer)`` into ONE pass on the GPU, so the raw matrix
        never exists and the values cross PCIe once (``return_trained=True`` also returns the trained standardizer);
        ``out=`` an existing [iid_count, sid_count] array of the right dtype / order to fill instead of allocating -- with a
        page-locked one from :func:`pysnptools_b200.util.pinned_empty` the result arrives at PCIe speed (no staging copy).
        """
        dtype = np.dtype(dtype)
        plain = standardizer is None or isinstance(standardizer, Identity)
        spec = None if plain else standardizer._device_spec()
        if plain or spec is None or not self._can_fuse():
            # no fused decode + standardize for this reader / standardizer (in-memory data, DistributedBed, a mapped file, DiagKtoN ...):
            # read, then standardize the fresh copy in place -- what the reference does (snpreader.py:419-478 + snpdata.py:138-188)
            takes_out = out is not None and self._can_fuse()
            val = self._read(None, None, order, dtype, force_python_only, view_ok and plain, num_threads, to_device=to_device,
                             **({"_out": out} if takes_out else {}))
            if out is not None and not takes_out:
                if _is_tensor(val) or not (isinstance(out, np.ndarray) and out.shape == val.shape and out.dtype == val.dtype and out.flags["WRITEABLE"]):
                    raise ValueError("out= must be a writeable ndarray of shape {0} and dtype {1}".format(tuple(val.shape), dtype))
                out[...] = val
                val = out
            data = SnpData(self.iid, self.sid, val, pos=self.pos, name=str(self), _require_float32_64=_require_float32_64)
            if plain:
                return (data, standardizer) if return_trained else data
            data, trained = data.standardize(standardizer, return_trained=True, force_python_only=force_python_only, num_threads=num_threads)
            return (data, trained) if return_trained else data
        stats_in = standardizer._trained_stats_for(self.sid)
        kw = {} if out is None else {"_out": out}
        val, stats = self._read(None, None, order, dtype, force_python_only, view_ok, num_threads, to_device=to_device,
                                _standardize=(spec, stats_in), **kw)
        data = SnpData(self.iid, self.sid, val, pos=self.pos, name=str(self))
        data._std_string_list.append(str(standardizer))
        if return_trained:
            return data, standardizer._make_trained(self.sid, np.asarray(stats, dtype=dtype))
        return data

    def _can_fuse(self):
        """True when ``_read`` accepts ``_standardize=`` / ``_out=`` (a packed .bed store behind it): decode + standardize in one GPU
        pass.  Readers without one (SnpData, DistributedBed, SnpMemMap) read first and standardize the copy."""
        return False

    def read_kernel(self, standardizer=None, block_size=None, order="A", dtype=np.float64, force_python_only=False,
                    view_ok=False, num_threads=None):
        """``K = X X^T`` of the standardized SNPs as a :class:`KernelData` (snpreader.py:528-561)."""
        assert standardizer is not None, "'standardizer' must be provided"
        from .kernelreader import SnpKernel
        return SnpKernel(self, standardizer=standardizer, block_size=block_size).read(
            order=order, dtype=dtype, force_python_only=force_python_only, view_ok=view_ok, num_threads=num_threads)

    def kernel(self, standardizer, allowlowrank=False, block_size=10000, blocksize=None, num_threads=None):
        """Deprecated spelling kept by the reference; returns the ndarray."""
        warnings.warn(".kernel(...) is deprecated. Use '.read_kernel(...).val", DeprecationWarning)
        return self.read_kernel(standardizer, block_size=blocksize or block_size).val

    def _read_kernel(self, standardizer, block_size=None, order="A", dtype=np.float64, force_python_only=False, view_ok=False,
                     return_trained=False, num_threads=None, to_device=False):
        """The block loop of snpreader.py:623-668 as ONE fused GPU pass: decode + standardize + tcgen05 SYRK per SNP chunk."""
        _no_python_path(force_python_only)
        from . import device
        dtype = np.dtype(dtype)
        try:
            root, iid_idx, sid_idx = self._root_and_indices()
        except NotImplementedError:
            # a subset of in-memory / memory-mapped values (e.g. what util.intersect_apply builds): materialise it and take the
            # float-matrix kernel (snpdata.py:190-214), as the reference does through SnpReader._as_snpdata
            data = self.read(order="A", dtype=dtype if dtype in (np.float32, np.float64) else np.float64, view_ok=True, num_threads=num_threads)
            return data._read_kernel(standardizer, block_size=block_size, order=order, dtype=dtype, force_python_only=force_python_only,
                                     view_ok=view_ok, return_trained=return_trained, num_threads=num_threads, to_device=to_device)
        if hasattr(root, "_read_kernel_pieces"):                  # DistributedBed: one packed store per piece, K accumulated
            if not isinstance(standardizer, Standardizer) or (standardizer._device_spec() is None and not isinstance(standardizer, Identity)):
                raise NotImplementedError("read_kernel on the GPU supports Unit, Beta, their trained forms and Identity")
            res = root._read_kernel_pieces(iid_idx, sid_idx, standardizer, block_size, dtype, return_trained)
            val = res[0] if return_trained else res
            val = np.asarray(val, order="F" if order == "F" else "C")
            return (val, res[1]) if return_trained else val
        sid_labels = self.sid
        spec = standardizer._device_spec() if isinstance(standardizer, Standardizer) else None
        stats = standardizer._trained_stats_for(sid_labels) if isinstance(standardizer, Standardizer) else None
        if spec is None:
            if not isinstance(standardizer, Identity):
                raise NotImplementedError("read_kernel on the GPU supports Unit, Beta, their trained forms and Identity")
            # Identity: x = dosage, missing -> 0 (the reference would propagate NaN; documented difference)
            spec, stats = ("unit",), np.tile(np.array([[0.0, 1.0]]), (len(sid_labels), 1))
        chunk = _kernel_chunk(block_size, self.iid_count, self.sid_count)
        exact = dtype == np.float64 and _KERNEL_FLOAT64[0] == "exact"
        if not to_device and getattr(root, "_device_store", True) is None and dtype in (np.float32, np.float64):
            # file -> host K in one call: the packed records are streamed to the GPU while earlier ones are multiplied
            val, st = root._kernel_host(iid_idx, sid_idx, spec, stats, dtype, chunk, exact=exact)
            if order == "F":
                val = val.T
            return (val, standardizer._make_trained(sid_labels, st.astype(dtype))) if return_trained else val
        store, ssel_local = root._store_for(sid_idx)
        if exact:
            out, d_stats = device.snp_kernel_f64(store, iid_idx, ssel_local, count_A1=root.count_A1, standardizer=spec, stats=stats)
        else:
            K32, d_stats = device.snp_kernel(store, iid_idx, ssel_local, count_A1=root.count_A1, standardizer=spec, stats=stats, chunk=chunk)
            out = device.convert_kernel(K32, dtype)
        val = out if to_device else _symmetric_to_host(out, order)
        if return_trained:
            st = d_stats.cpu().numpy().astype(dtype if dtype in (np.float32, np.float64) else np.float64)
            return val, standardizer._make_trained(sid_labels, st)
        return val

    def _root_and_indices(self):
        raise NotImplementedError

    def __repr__(self):
        return "{0}()".format(self.__class__.__name__)


_KERNEL_FLOAT64 = ["tensor"]


def set_kernel_float64(mode):
    """What a ``dtype=float64`` kernel request computes with: ``"tensor"`` (default) -- the tcgen05 path (fp32 accumulation, <= 1e-5
    relative Frobenius error, the north_star gate) converted to float64; ``"exact"`` -- float64 arithmetic on the GPU
    (``pstb_snp_kernel_f64`` / ``pstb_float_kernel_f64``), which is what the reference's ``val.dot(val.T)`` is for float64 values
    (snpdata.py:203-206) and what its 10-decimal unit tests need.  ``dtype=float32`` requests always use the tensor cores.
    Returns the previous mode."""
    if mode not in ("tensor", "exact"):
        raise ValueError("mode must be 'tensor' or 'exact'")
    prev = _KERNEL_FLOAT64[0]
    _KERNEL_FLOAT64[0] = mode
    return prev


def _kernel_chunk(block_size, n_iid, n_sid):
    from . import device
    chunk = device.default_kernel_chunk(n_iid, n_sid)
    if block_size is not None:
        chunk = min(chunk, max(64, (int(block_size) + 63) // 64 * 64))
    return chunk


class _SnpSubset(SnpReader):
    """Lazy ``reader[iid_indexer, sid_indexer]`` (snpreader/_subset.py, pstreader/_subset.py)."""

    def __init__(self, internal, iid_indexer, sid_indexer):
        self._internal = internal
        self._iid_index = _resolve_indexer(iid_indexer, internal.iid_count)
        self._sid_index = _resolve_indexer(sid_indexer, internal.sid_count)

    def __repr__(self):
        def fmt(ix):
            return ":" if ix is None else ("[" + ",".join(str(i) for i in ix[:8]) + (",..." if len(ix) > 8 else "") + "]")
        return "{0}[{1},{2}]".format(self._internal, fmt(self._iid_index), fmt(self._sid_index))

    @property
    def row(self):
        r = self._internal.row
        return r if self._iid_index is None else r[self._iid_index]

    @property
    def col(self):
        c = self._internal.col
        return c if self._sid_index is None else c[self._sid_index]

    @property
    def col_property(self):
        c = self._internal.col_property
        return c if self._sid_index is None else c[self._sid_index]

    @property
    def iid_count(self):
        return self._internal.iid_count if self._iid_index is None else len(self._iid_index)

    @property
    def sid_count(self):
        return self._internal.sid_count if self._sid_index is None else len(self._sid_index)

    def _read(self, iid_index_or_none, sid_index_or_none, order, dtype, force_python_only, view_ok, num_threads, to_device=False,
              _standardize=None, _out=None):
        kw = {} if _standardize is None else {"_standardize": _standardize}
        if _out is not None:
            kw["_out"] = _out
        return self._internal._read(_compose(self._iid_index, iid_index_or_none), _compose(self._sid_index, sid_index_or_none),
                                    order, dtype, force_python_only, view_ok, num_threads, to_device=to_device, **kw)

    def _can_fuse(self):
        return self._internal._can_fuse()

    def _root_and_indices(self):
        root, ii, si = self._internal._root_and_indices()
        return root, _compose(ii, self._iid_index), _compose(si, self._sid_index)


class Bed(SnpReader):
    """PLINK ``.bed/.bim/.fam`` reader decoded on the GPU (reference: snpreader/bed.py)."""

    def __init__(self, filename, count_A1=None, iid=None, sid=None, pos=None, num_threads=None, skip_format_check=False,
                 fam_filename=None, bim_filename=None, chrom_map=plink_chrom_map):
        super(Bed, self).__init__()
        filename = str(filename)
        self.filename = filename if filename.endswith(".bed") else filename + ".bed"
        self.fam_filename = fam_filename or self.filename[:-4] + ".fam"
        self.bim_filename = bim_filename or self.filename[:-4] + ".bim"
        if count_A1 is None:
            warnings.warn("'count_A1' was not set. For now it will default to 'False', but in the future it will default to 'True'", FutureWarning)
            count_A1 = False
        self.count_A1 = count_A1
        self._skip_format_check = skip_format_check
        self._original_iid, self._original_sid, self._original_pos = iid, sid, pos
        self._num_threads = num_threads
        self.chrom_map = chrom_map
        self._row = self._col = self._col_property = None
        self._host_packed = None
        self._device_store = None

    def __repr__(self):
        return "{0}('{1}',count_A1={2})".format(self.__class__.__name__, self.filename, self.count_A1)

    def __getstate__(self):                       # survives (cloud)pickle: drop file / device handles (test.py:993-1003)
        d = dict(self.__dict__)
        d["_host_packed"] = None
        d["_device_store"] = None
        return d

    # --- metadata (bed.py:147-194) ---
    @staticmethod
    def _read_columns(path, cols):
        out = [[] for _ in cols]
        with open(path) as f:
            for line in f:
                parts = line.split()
                if not parts:
                    continue
                for k, c in enumerate(cols):
                    out[k].append(parts[c])
        return out

    @property
    def row(self):
        if self._row is None:
            if self._original_iid is not None:
                self._row = np.array(self._original_iid, dtype=str).reshape(-1, 2)
            else:
                fid, iid = self._read_columns(self.fam_filename, (0, 1))
                self._row = np.array([fid, iid], dtype=str).T.reshape(-1, 2)
        return self._row

    @property
    def col(self):
        if self._col is None:
            self._load_bim()
        return self._col

    @property
    def col_property(self):
        if self._col_property is None:
            self._load_bim()
        return self._col_property

    def _load_bim(self):
        need_file = self._original_sid is None or self._original_pos is None
        chrom = sid = cm = bp = None
        if need_file:
            chrom, sid, cm, bp = self._read_columns(self.bim_filename, (0, 1, 2, 3))
        self._col = np.array(self._original_sid if self._original_sid is not None else sid, dtype=str)
        if self._original_pos is not None:
            pos = np.array(self._original_pos, dtype=np.float64).reshape(-1, 3)
        else:
            def chrom_value(c):
                if c in self.chrom_map:
                    return float(self.chrom_map[c])
                try:
                    return float(c)
                except ValueError:
                    raise ValueError("chromosome '{0}' in '{1}' is not a number or one of {2}".format(
                        c, self.bim_filename, sorted(self.chrom_map)))
            pos = np.array([[chrom_value(c) for c in chrom], [float(x) for x in cm], [float(x) for x in bp]], dtype=np.float64).T.reshape(-1, 3)
            pos[pos == 0] = np.nan
        self._col_property = pos
        if len(self._col_property) != len(self._col):
            raise ValueError("pos and sid must have the same length")

    # --- packed bytes ---
    def _packed_host(self):
        """File bytes after the 3-byte header as a read-only uint8 [sid_count, ceil(iid_count/4)] memory map."""
        if self._host_packed is None:
            n, m = self.iid_count, self.sid_count
            rec = (n + 3) // 4
            size = os.path.getsize(self.filename)
            with open(self.filename, "rb") as f:
                head = f.read(3)
            if not self._skip_format_check and head != bytes([0x6C, 0x1B, 0x01]):
                raise ValueError("'{0}' is not a SNP-major PLINK .bed file (bad magic bytes)".format(self.filename))
            if size != 3 + m * rec:
                raise ValueError("'{0}': expected {1} bytes for {2} iids x {3} sids, found {4}".format(self.filename, 3 + m * rec, n, m, size))
            self._host_packed = np.memmap(self.filename, dtype=np.uint8, mode="r", offset=3, shape=(m, rec)) if m * rec else np.zeros((m, rec), np.uint8)
        return self._host_packed

    def _store_for(self, sid_idx):
        """Packed store in HBM holding (at least) the requested SNP records + the selection in store coordinates."""
        from . import device
        if self._device_store is None:
            self._device_store = device.PackedStore.from_host(np.asarray(self._packed_host()), self.iid_count)
        return self._device_store, sid_idx

    def _root_and_indices(self):
        return self, None, None

    def _can_fuse(self):
        return True

    def _kernel_host(self, iid_idx, sid_idx, spec, stats_in, dtype, chunk, exact=False):
        """``pstb_snp_kernel_host``: memory-mapped file bytes -> K as a NumPy array (C order, symmetric) + float64 statistics."""
        packed = self._packed_host()
        n, m = self.iid_count, self.sid_count
        ii = None if iid_idx is None else np.ascontiguousarray(iid_idx, dtype=np.int64)
        si = None if sid_idx is None else np.ascontiguousarray(sid_idx, dtype=np.int64)
        ni, ns = (n if ii is None else len(ii)), (m if si is None else len(si))
        K = np.empty((ni, ni), dtype=dtype)
        stats = np.empty((ns, 2), dtype=np.float64)
        use_stats = 0
        if stats_in is not None:
            stats[...] = np.asarray(stats_in, dtype=np.float64)
            use_stats = 1
        mode = _lib.STD_UNIT if spec[0] == "unit" else _lib.STD_BETA
        a, b = (float(spec[1]), float(spec[2])) if spec[0] == "beta" else (float("nan"), float("nan"))
        if ni:
            _lib.require_gpu()
            args = (packed.ctypes.data if m else None, n, m, ii.ctypes.data if ii is not None else None, ni,
                    si.ctypes.data if si is not None else None, ns, int(bool(self.count_A1)), mode, a, b, use_stats, stats.ctypes.data, K.ctypes.data)
            if exact:
                _lib.check(_lib.lib.pstb_snp_kernel_host_f64(*args, int(chunk)))
            else:
                _lib.check(_lib.lib.pstb_snp_kernel_host(*args, _DT_CODE[np.dtype(dtype)], int(chunk), _lib.LOW_TERM_DEFAULT))
        return K, stats

    # --- read (bed.py:318-345 -> bed_reader read_f32/f64/i8) ---
    def _read(self, iid_index_or_none, sid_index_or_none, order, dtype, force_python_only, view_ok, num_threads, to_device=False,
              _standardize=None, _out=None):
        _no_python_path(force_python_only)
        dtype = np.dtype(dtype)
        if dtype not in _DT_CODE:
            raise ValueError("dtype must be float32, float64 or int8")
        if order == "A":
            order = "F"
        spec, stats_in = _standardize if _standardize is not None else (None, None)
        if spec is not None and dtype == np.int8:
            raise ValueError("standardizing needs a float32 / float64 read")
        if to_device:
            from . import device
            store, ssel = self._store_for(sid_index_or_none)
            val, st = device.read(store, iid_index_or_none, ssel, count_A1=self.count_A1, dtype=dtype, order=order,
                                  standardizer=spec, stats=stats_in)
            return val if spec is None else (val, st.cpu().numpy())
        packed = self._packed_host()
        n, m = self.iid_count, self.sid_count
        ii = None if iid_index_or_none is None else np.ascontiguousarray(iid_index_or_none, dtype=np.int64)
        si = None if sid_index_or_none is None else np.ascontiguousarray(sid_index_or_none, dtype=np.int64)
        for idx, cnt in ((ii, n), (si, m)):
            if idx is not None and idx.size and (idx.min() < 0 or idx.max() >= cnt):
                raise IndexError("index out of range for axis of size {0}".format(cnt))
        ni, ns = (n if ii is None else len(ii)), (m if si is None else len(si))
        if _out is not None:
            want = "F_CONTIGUOUS" if order == "F" else "C_CONTIGUOUS"
            if not (isinstance(_out, np.ndarray) and _out.shape == (ni, ns) and _out.dtype == dtype and _out.flags[want] and _out.flags["WRITEABLE"]):
                raise ValueError("out= must be a writable {0} ndarray of shape ({1}, {2}) in order '{3}'".format(dtype, ni, ns, order))
            val = _out
        else:
            val = np.empty((ni, ns), dtype=dtype, order=order)
        mode, a, b, use_stats = _lib.STD_NONE, 0.0, 0.0, 0
        stats = None
        if spec is not None:
            mode = _lib.STD_UNIT if spec[0] == "unit" else _lib.STD_BETA
            a, b = (float(spec[1]), float(spec[2])) if spec[0] == "beta" else (float("nan"), float("nan"))
            stats = np.empty((ns, 2), dtype=np.float64)
            if stats_in is not None:
                stats[...] = np.asarray(stats_in, dtype=np.float64)
                use_stats = 1
        if ni and ns:
            _lib.require_gpu()
            _lib.check(_lib.lib.pstb_read_host(packed.ctypes.data, n, m, ii.ctypes.data if ii is not None else None, ni,
                                               si.ctypes.data if si is not None else None, ns, int(bool(self.count_A1)),
                                               mode, a, b, use_stats, stats.ctypes.data if stats is not None else None,
                                               val.ctypes.data, _DT_CODE[dtype], _order_code(order)))
        return val if spec is None else (val, stats)

    def copyinputs(self, copier):
        """bed.py:196-201: the three files this reader needs."""
        copier.input(self.filename)
        copier.input(self.fam_filename)
        copier.input(self.bim_filename)

    # --- write (bed.py:229-316 -> to_bed) ---
    @staticmethod
    def write(filename, snpdata, count_A1=False, force_python_only=False, _require_float32_64=True, num_threads=None,
              reverse_chrom_map={}):
        """Pack ``snpdata.val`` ({0,1,2,NaN} or int8 with -127) on the GPU and write ``.bed/.fam/.bim``; returns a Bed.
        ``reverse_chrom_map`` (e.g. ``{23: 'X'}``) maps chromosome numbers back to names in the ``.bim`` (bed.py:293-298)."""
        _no_python_path(force_python_only)
        if isinstance(filename, SnpReader) and isinstance(snpdata, str):        # historical argument order (bed.py:275-282)
            warnings.warn("write statement should have filename before data to write", DeprecationWarning)
            filename, snpdata = snpdata, filename
        import torch
        from . import device
        filename = str(filename)
        filename = filename if filename.endswith(".bed") else filename + ".bed"
        if isinstance(snpdata, SnpReader) and not hasattr(snpdata, "val"):
            snpdata = snpdata.read(dtype=np.float64 if _require_float32_64 else np.int8, _require_float32_64=_require_float32_64)
        val = snpdata.val
        n, m = val.shape
        rec = (n + 3) // 4
        if n and m:
            t = val if _is_tensor(val) else torch.from_numpy(np.ascontiguousarray(val)).cuda()
            packed = device.pack(t, count_A1=count_A1).tensor[:, :rec].contiguous().cpu().numpy()
        else:
            packed = np.zeros((m, rec), dtype=np.uint8)
        with open(filename, "wb") as f:
            f.write(bytes([0x6C, 0x1B, 0x01]))
            f.write(packed.tobytes())
        with open(filename[:-4] + ".fam", "w") as f:
            for fid, iid in snpdata.iid:
                f.write("{0} {1} 0 0 0 0\n".format(fid, iid))
        pos = snpdata.pos
        with open(filename[:-4] + ".bim", "w") as f:
            for k, sid in enumerate(snpdata.sid):
                c, cm, bp = (0 if np.isnan(x) else x for x in pos[k])
                c = reverse_chrom_map.get(c, int(c) if float(c).is_integer() else c)
                f.write("{0}\t{1}\t{2}\t{3}\tA\tC\n".format(c, sid, cm, int(bp)))
        return Bed(filename, count_A1=count_A1)


class SnpData(SnpReader):
    """In-memory SNP values + labels (reference: snpreader/snpdata.py).  ``val`` is a NumPy array or a CUDA tensor."""

    def __init__(self, iid, sid, val, pos=None, name=None, parent_string=None, copyinputs_function=None, xp=None,
                 _require_float32_64=True):
        super(SnpData, self).__init__()
        self._row = np.array(iid, dtype=str).reshape(-1, 2)
        self._col = np.array(sid, dtype=str).reshape(-1)
        if pos is None:
            pos = np.full((len(self._col), 3), np.nan)
        self._col_property = np.array(pos, dtype=np.float64).reshape(-1, 3)
        if not _is_tensor(val):
            val = val if isinstance(val, np.ndarray) else np.array(val, dtype=np.float64)       # an np.memmap stays one (view_ok reads)
            if _require_float32_64 and val.dtype not in (np.float32, np.float64):
                val = val.astype(np.float64)
            if val.ndim != 2:
                val = val.reshape(len(self._row), len(self._col))
        assert tuple(val.shape) == (len(self._row), len(self._col)), "val must be [iid_count, sid_count]"
        self._val = val
        self._name = name or parent_string or ""
        self._std_string_list = []

    @property
    def val(self):
        return self._val

    @val.setter
    def val(self, new_value):
        self._val = new_value

    @property
    def row(self):
        return self._row

    @property
    def col(self):
        return self._col

    @property
    def col_property(self):
        return self._col_property

    def __repr__(self):
        if self._name == "":
            s = "SnpData()" if not self._std_string_list else "SnpData({0})".format(",".join(self._std_string_list))
        else:
            s = "SnpData({0})".format(",".join([self._name] + self._std_string_list))
        return s

    def _read(self, iid_index_or_none, sid_index_or_none, order, dtype, force_python_only, view_ok, num_threads, to_device=False):
        """Sub-matrix of the in-memory values (pstdata.py:217-222 -> util.sub_matrix -> pstb_subset_host)."""
        val = self._val
        dtype = np.dtype(dtype)
        if _is_tensor(val):
            import torch
            out = val
            if iid_index_or_none is not None:
                out = out[torch.as_tensor(iid_index_or_none, device=val.device)]
            if sid_index_or_none is not None:
                out = out[:, torch.as_tensor(sid_index_or_none, device=val.device)]
            return out if to_device else np.asarray(out.cpu().numpy().astype(dtype), order="F" if order in ("F", "A") else "C")
        if iid_index_or_none is None and sid_index_or_none is None:
            ok_order = order == "A" or val.flags["F_CONTIGUOUS" if order == "F" else "C_CONTIGUOUS"]
            if view_ok and ok_order and val.dtype == dtype:
                return val
            return np.array(val, dtype=dtype, order=order)
        from .util import sub_matrix
        ii = np.arange(val.shape[0]) if iid_index_or_none is None else iid_index_or_none
        si = np.arange(val.shape[1]) if sid_index_or_none is None else sid_index_or_none
        return sub_matrix(val, ii, si, order="F" if order == "A" else order, dtype=dtype)

    def _root_and_indices(self):
        raise NotImplementedError("SnpData has no packed store")

    def allclose(self, value, equal_nan=True):
        """Same labels and close values (snpdata.py:108-124 -> pstdata.py allclose)."""
        def arr(v):
            return v.cpu().numpy() if _is_tensor(v) else np.asarray(v)
        return (np.array_equal(self.iid, value.iid) and np.array_equal(self.sid, value.sid)
                and np.allclose(self.pos, value.pos, equal_nan=True)
                and self.val.shape == value.val.shape and bool(np.allclose(arr(self.val), arr(value.val), equal_nan=equal_nan)))

    def train_standardizer(self, apply_in_place, standardizer=Unit(), force_python_only=False, num_threads=None):
        """Deprecated spelling (snpdata.py:126-135): ``standardize(..., return_trained=True)``."""
        warnings.warn("train_standardizer is deprecated. standardize(...,return_trained=True,...) instead", DeprecationWarning)
        assert apply_in_place, "code assumes apply_in_place"
        return self.standardize(standardizer, return_trained=True, force_python_only=force_python_only, num_threads=num_threads)[1]

    def standardize(self, standardizer=Unit(), block_size=None, return_trained=False, force_python_only=False, num_threads=None):
        """In-place standardize; returns self (and the trained standardizer) -- snpdata.py:138-188."""
        self._std_string_list.append(str(standardizer))
        _, trained = standardizer.standardize(self, return_trained=True, force_python_only=force_python_only, num_threads=num_threads)
        return (self, trained) if return_trained else self

    def _read_kernel(self, standardizer, block_size=None, order="A", dtype=np.float64, force_python_only=False, view_ok=False,
                     return_trained=False, num_threads=None, to_device=False):
        """``val.dot(val.T)`` (snpdata.py:190-214) on the tensor cores via fp16 hi/lo operand planes."""
        _no_python_path(force_python_only)
        from . import device
        import torch
        dtype = np.dtype(dtype)
        if isinstance(standardizer, Identity):
            data, trained = self, standardizer
        else:
            data = SnpData(self.iid, self.sid, self.val.clone() if _is_tensor(self.val) else np.array(self.val, order="A"), pos=self.pos)
            data, trained = data.standardize(standardizer, return_trained=True, num_threads=num_threads)
        v = data.val if _is_tensor(data.val) else torch.from_numpy(np.ascontiguousarray(data.val)).cuda()
        if dtype == np.float64 and _KERNEL_FLOAT64[0] == "exact":
            out = device.float_kernel_f64(v.double())
        else:
            K32 = device.float_kernel(v)
            out = device.convert_kernel(K32, dtype)
        val = out if to_device else _symmetric_to_host(out, order)
        return (val, trained) if return_trained else val


def __getattr__(name):
    """``pysnptools_b200.snpreader.DistributedBed`` as in the reference's package layout (imported lazily: it builds on this module)."""
    if name == "DistributedBed":
        from .distributedbed import DistributedBed
        return DistributedBed
    if name == "SnpMemMap":
        from .snpmemmap import SnpMemMap
        return SnpMemMap
    raise AttributeError("module {0!r} has no attribute {1!r}".format(__name__, name))
