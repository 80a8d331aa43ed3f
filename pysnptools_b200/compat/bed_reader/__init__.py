"""Drop-in for the nine ``bed_reader`` symbols PySnpTools imports, backed by ``libpst_b200.so``.

Put ``pysnptools_b200/compat`` on ``PYTHONPATH`` and the unmodified reference package
(``pysnptools/snpreader/bed.py:5``, ``standardizer/standardizer.py:4``, ``util/__init__.py:12``) runs its
decode / standardize / gather on the GPU.  See INTEGRATION.md.
"""
import os

import numpy as np

from pysnptools_b200 import _lib
from pysnptools_b200.util import get_num_threads  # noqa: F401  (re-exported symbol)

_DT = {np.dtype(np.float32): _lib.F32, np.dtype(np.float64): _lib.F64, np.dtype(np.int8): _lib.I8}
_FAM = {"fid": 0, "iid": 1, "father": 2, "mother": 3, "sex": 4, "pheno": 5}
_BIM = {"chromosome": 0, "sid": 1, "cm_position": 2, "bp_position": 3, "allele_1": 4, "allele_2": 5}
_KIND = {"cm_position": np.float64, "bp_position": np.int64, "sex": np.int32}


class open_bed(object):
    """``bed_reader.open_bed`` as PySnpTools uses it (bed.py:137-145, 337-343; snpreader.py:720,735)."""

    def __init__(self, filepath, iid_count=None, sid_count=None, properties={}, count_A1=True, num_threads=None,
                 skip_format_check=False, fam_filepath=None, bim_filepath=None):
        self.filepath = str(filepath)
        base = self.filepath[:-4] if self.filepath.endswith(".bed") else self.filepath
        self.fam_filepath = str(fam_filepath) if fam_filepath is not None else base + ".fam"
        self.bim_filepath = str(bim_filepath) if bim_filepath is not None else base + ".bim"
        self.count_A1 = count_A1
        self._num_threads = num_threads
        self._skip_format_check = skip_format_check
        self._properties = dict(properties)
        self._cache = {}
        self._iid_count, self._sid_count = iid_count, sid_count
        self._packed = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self._packed = None
        return False

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_packed"] = None
        return d

    def _column(self, name):
        if name not in self._cache:
            given = self._properties.get(name, "unset")
            if given is None:
                raise AttributeError("property '{0}' was skipped (properties[{0!r}] = None)".format(name))
            if not isinstance(given, str):
                val = np.asarray(given)
            else:
                table, path = (_FAM, self.fam_filepath) if name in _FAM else (_BIM, self.bim_filepath)
                col = table[name]
                with open(path) as f:
                    items = [parts[col] for parts in (line.split() for line in f) if parts]
                val = np.array(items, dtype=_KIND.get(name, str)) if name not in ("bp_position",) else np.array(items, dtype=np.float64).astype(np.int64)
            self._cache[name] = val
        return self._cache[name]

    fid = property(lambda self: self._column("fid"))
    iid = property(lambda self: self._column("iid"))
    father = property(lambda self: self._column("father"))
    mother = property(lambda self: self._column("mother"))
    sex = property(lambda self: self._column("sex"))
    pheno = property(lambda self: self._column("pheno"))
    chromosome = property(lambda self: self._column("chromosome"))
    sid = property(lambda self: self._column("sid"))
    cm_position = property(lambda self: self._column("cm_position"))
    bp_position = property(lambda self: self._column("bp_position"))
    allele_1 = property(lambda self: self._column("allele_1"))
    allele_2 = property(lambda self: self._column("allele_2"))

    @property
    def iid_count(self):
        if self._iid_count is None:
            self._iid_count = len(self.iid)
        return self._iid_count

    @property
    def sid_count(self):
        if self._sid_count is None:
            self._sid_count = len(self.sid)
        return self._sid_count

    @property
    def shape(self):
        return (self.iid_count, self.sid_count)

    def _packed_host(self):
        if self._packed is None:
            n, m = self.iid_count, self.sid_count
            rec = (n + 3) // 4
            with open(self.filepath, "rb") as f:
                head = f.read(3)
            if not self._skip_format_check and head != bytes([0x6C, 0x1B, 0x01]):
                raise ValueError("'{0}' is not a SNP-major PLINK .bed file (bad magic bytes)".format(self.filepath))
            if os.path.getsize(self.filepath) != 3 + m * rec:
                raise ValueError("'{0}' has the wrong size for {1} x {2}".format(self.filepath, n, m))
            self._packed = np.memmap(self.filepath, dtype=np.uint8, mode="r", offset=3, shape=(m, rec)) if m * rec else np.zeros((m, rec), np.uint8)
        return self._packed

    @staticmethod
    def _index(ix, count):
        if ix is None:
            return None
        if isinstance(ix, slice):
            return np.arange(count, dtype=np.int64)[ix]
        a = np.asarray(ix)
        if a.dtype == bool:
            return np.nonzero(a)[0].astype(np.int64)
        a = np.atleast_1d(a).astype(np.int64)
        a = np.where(a < 0, a + count, a)
        if a.size and (a.min() < 0 or a.max() >= count):
            raise IndexError("index out of range for axis of size {0}".format(count))
        return np.ascontiguousarray(a)

    def read(self, index=None, dtype="float32", order="F", force_python_only=False, num_threads=None):
        if force_python_only:
            raise NotImplementedError("the CUDA bed_reader has no pure-Python path")
        if not isinstance(index, tuple):
            index = (None, index)
        n, m = self.iid_count, self.sid_count
        ii, si = self._index(index[0], n), self._index(index[1], m)
        dtype = np.dtype(dtype)
        ni, ns = (n if ii is None else len(ii)), (m if si is None else len(si))
        val = np.empty((ni, ns), dtype=dtype, order=order)
        if ni and ns:
            packed = self._packed_host()
            _lib.require_gpu()
            _lib.check(_lib.lib.pstb_read_host(packed.ctypes.data, n, m, ii.ctypes.data if ii is not None else None, ni,
                                               si.ctypes.data if si is not None else None, ns, int(bool(self.count_A1)), _lib.STD_NONE,
                                               0.0, 0.0, 0, None, val.ctypes.data, _DT[dtype], _lib.ORDER_C if order == "C" else _lib.ORDER_F))
        return val


def _require_writeable(a):
    """The C ABI writes through raw pointers: refuse read-only arrays (a memory map opened with mode='r' would take the process
    down; the Rust extension raises for non-writeable arrays as well)."""
    if not a.flags.writeable:
        raise ValueError("assignment destination is read-only")


def _standardize(snps, is_beta, a, b, apply_in_place, use_stats, stats, num_threads, code):
    if apply_in_place:
        _require_writeable(snps)
    if not use_stats:
        _require_writeable(stats)
    n_iid, n_sid = snps.shape
    st64 = np.array(stats, dtype=np.float64, order="C") if use_stats else np.empty((n_sid, 2), dtype=np.float64)
    order = _lib.ORDER_C if snps.flags["C_CONTIGUOUS"] else _lib.ORDER_F
    if n_sid:
        _lib.require_gpu()
        _lib.check(_lib.lib.pstb_standardize_host(snps.ctypes.data, code, order, n_iid, n_sid, _lib.STD_BETA if is_beta else _lib.STD_UNIT,
                                                  float(a), float(b), int(bool(apply_in_place)), int(bool(use_stats)), st64.ctypes.data))
    if not use_stats:
        stats[...] = st64


def standardize_f64(snps, is_beta, a, b, apply_in_place, use_stats, stats, num_threads):
    _standardize(snps, is_beta, a, b, apply_in_place, use_stats, stats, num_threads, _lib.F64)


def standardize_f32(snps, is_beta, a, b, apply_in_place, use_stats, stats, num_threads):
    _standardize(snps, is_beta, a, b, apply_in_place, use_stats, stats, num_threads, _lib.F32)


def _subset(val, rows, cols, out, num_threads, ci, co):
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    cols = np.ascontiguousarray(cols, dtype=np.int64)
    _require_writeable(out)
    if out.size:
        _lib.require_gpu()
        _lib.check(_lib.lib.pstb_subset_host(val.ctypes.data, ci, _lib.ORDER_C if val.flags["C_CONTIGUOUS"] else _lib.ORDER_F, val.shape[0],
                                             val.shape[1], val.shape[2], rows.ctypes.data, len(rows), cols.ctypes.data, len(cols),
                                             out.ctypes.data, co, _lib.ORDER_C if out.flags["C_CONTIGUOUS"] else _lib.ORDER_F))


def subset_f64_f64(val, rows, cols, out, num_threads):
    _subset(val, rows, cols, out, num_threads, _lib.F64, _lib.F64)


def subset_f32_f64(val, rows, cols, out, num_threads):
    _subset(val, rows, cols, out, num_threads, _lib.F32, _lib.F64)


def subset_f32_f32(val, rows, cols, out, num_threads):
    _subset(val, rows, cols, out, num_threads, _lib.F32, _lib.F32)


def to_bed(filepath, val, properties={}, count_A1=True, fam_filepath=None, bim_filepath=None, force_python_only=False, num_threads=None):
    """Write ``val`` ({0,1,2,NaN} floats or int8 with -127) as .bed/.fam/.bim, packing on the GPU (bed.py:300-314)."""
    if force_python_only:
        raise NotImplementedError("the CUDA bed_reader has no pure-Python path")
    import torch
    from pysnptools_b200 import device
    filepath = str(filepath)
    base = filepath[:-4] if filepath.endswith(".bed") else filepath
    val = np.asarray(val)
    n, m = val.shape
    rec = (n + 3) // 4
    if n and m:
        packed = device.pack(torch.from_numpy(np.ascontiguousarray(val)).cuda(), count_A1=count_A1).tensor[:, :rec].contiguous().cpu().numpy()
    else:
        packed = np.zeros((m, rec), dtype=np.uint8)
    with open(filepath, "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01]))
        f.write(packed.tobytes())

    def text(x, missing, integral):
        # NaN -> the column's missing value ('0'); integral floats without a decimal point (PLINK writes '1', not '1.0')
        if isinstance(x, (float, np.floating)):
            if x != x:
                return missing
            if integral or float(x).is_integer():
                return str(int(x))
            return repr(float(x))
        return str(x)

    def col(name, count, default, missing="0", integral=False):
        v = properties.get(name)
        if v is None:
            return [default(k) for k in range(count)]
        if len(v) != count:
            raise ValueError("property '{0}' has {1} entries, expected {2}".format(name, len(v), count))
        return [text(x, missing, integral) for x in list(v)]
    fam = [col("fid", n, lambda k: "0"), col("iid", n, lambda k: "iid{0}".format(k + 1)), col("father", n, lambda k: "0"),
           col("mother", n, lambda k: "0"), col("sex", n, lambda k: 0), col("pheno", n, lambda k: "0")]
    with open(fam_filepath or base + ".fam", "w") as f:
        for k in range(n):
            f.write(" ".join(str(c[k]) for c in fam) + "\n")
    bim = [col("chromosome", m, lambda k: "0"), col("sid", m, lambda k: "sid{0}".format(k + 1)), col("cm_position", m, lambda k: 0.0),
           col("bp_position", m, lambda k: 0, integral=True), col("allele_1", m, lambda k: "A1"), col("allele_2", m, lambda k: "A2")]
    with open(bim_filepath or base + ".bim", "w") as f:
        for k in range(m):
            f.write("\t".join(str(c[k]) for c in bim) + "\n")
