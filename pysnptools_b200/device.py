"""Device-resident API: packed genotype store in HBM + the kernels of ``libpst_b200.so``.

PyTorch is used only for device buffers and streams; every computation is a call through the C ABI.
The NumPy-facing classes (:mod:`pysnptools_b200.snpreader` ...) are thin wrappers over this module.
"""
import numpy as np
import torch

from . import _lib
from ._lib import Axis, lib, check

BED_MAGIC = bytes([0x6C, 0x1B, 0x01])
_LOW_TERM = {"default": -1, "fp16": 0, "fp8": 1, "auto": 2}
_LOW_TERM_NAME = {0: "fp16", 1: "fp8", 2: "auto"}
_DT = {np.dtype(np.float32): (_lib.F32, torch.float32), np.dtype(np.float64): (_lib.F64, torch.float64),
       np.dtype(np.int8): (_lib.I8, torch.int8)}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _order_code(order):
    if order in ("F", "A"):
        return _lib.ORDER_F
    if order == "C":
        return _lib.ORDER_C
    raise ValueError("order must be 'F', 'C' or 'A', not {0!r}".format(order))


def _mode_args(standardizer_spec):
    """(mode, a, b) from None | ('unit',) | ('beta', a, b)."""
    if standardizer_spec is None:
        return _lib.STD_NONE, 0.0, 0.0
    if standardizer_spec[0] == "unit":
        return _lib.STD_UNIT, float("nan"), float("nan")
    if standardizer_spec[0] == "beta":
        return _lib.STD_BETA, float(standardizer_spec[1]), float(standardizer_spec[2])
    raise ValueError("unknown standardizer spec {0!r}".format(standardizer_spec))


class Selection(object):
    """One axis of a read: ``None`` (all), a slice, or an integer vector -> ``pstb_axis`` (+ keep-alive)."""

    def __init__(self, sel, count, device):
        self.count = int(count)
        self.keep = None
        if sel is None:
            self.start, self.step, self.n = 0, 1, self.count
            return
        if isinstance(sel, slice):
            r = range(self.count)[sel]
            self.start, self.step, self.n = (r.start, r.step, len(r)) if len(r) else (0, 1, 0)
            return
        idx = np.asarray(sel)
        if idx.dtype == bool:
            idx = np.nonzero(idx)[0]
        idx = idx.astype(np.int64).reshape(-1)
        idx = np.where(idx < 0, idx + self.count, idx)
        if idx.size and (idx.min() < 0 or idx.max() >= self.count):
            raise IndexError("index out of range for axis of size {0}".format(self.count))
        self.n = int(idx.size)
        if self.n <= 1:
            self.start, self.step = (int(idx[0]) if self.n else 0), 1
            return
        d = np.diff(idx)
        if d[0] != 0 and np.all(d == d[0]):           # identity / range / constant stride stay implicit
            self.start, self.step = int(idx[0]), int(d[0])
            return
        self.start, self.step = 0, 1
        # int32 carries the same bits as the uint32 the C ABI reads (indices < 2**31)
        self.keep = torch.from_numpy(idx.astype(np.int32)).to(device)

    def axis(self):
        return Axis(self.keep.data_ptr() if self.keep is not None else None, self.start, self.step, self.n)

    def indices(self):
        if self.keep is not None:
            return self.keep.cpu().numpy().astype(np.int64)
        return self.start + self.step * np.arange(self.n, dtype=np.int64)


class PackedStore(object):
    """SNP-major 2-bit records in HBM: uint8 tensor ``[sid_count, ld]`` with ``ld`` a 16-byte multiple."""

    def __init__(self, tensor, iid_count, sid_count):
        self.tensor, self.iid_count, self.sid_count = tensor, int(iid_count), int(sid_count)
        self.ld = int(tensor.shape[1]) if tensor.dim() == 2 else int(lib.pstb_packed_ld(iid_count))

    @property
    def device(self):
        return self.tensor.device

    @staticmethod
    def from_host(packed, iid_count, device="cuda"):
        """``packed``: uint8 ndarray ``[sid_count, ceil(iid_count/4)]`` (file bytes after the 3-byte header)."""
        _lib.require_gpu()
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        if not packed.flags.writeable:
            packed = packed.copy()                      # torch.from_numpy wants a writable buffer (memory-mapped files are not)
        sid_count, rec = packed.shape if packed.ndim == 2 else (0, 0)
        assert rec == (iid_count + 3) // 4 or sid_count == 0
        ld = int(lib.pstb_packed_ld(iid_count))
        t = torch.zeros((sid_count, max(ld, 16)), dtype=torch.uint8, device=device)
        if sid_count and rec:
            t[:, :rec].copy_(torch.from_numpy(packed), non_blocking=False)
        return PackedStore(t, iid_count, sid_count)

    @staticmethod
    def from_file(path, iid_count, sid_count, skip_format_check=False, device="cuda"):
        rec = (int(iid_count) + 3) // 4
        with open(path, "rb") as f:
            head = f.read(3)
            if not skip_format_check and head != BED_MAGIC:
                raise ValueError("'{0}' is not a SNP-major PLINK .bed file (bad magic bytes)".format(path))
            raw = np.fromfile(f, dtype=np.uint8)
        if raw.size != sid_count * rec:
            raise ValueError("'{0}': expected {1} genotype bytes for {2} x {3}, found {4}".format(
                path, sid_count * rec, iid_count, sid_count, raw.size))
        return PackedStore.from_host(raw.reshape(sid_count, rec), iid_count, device=device)


def _alloc_out(n_iid, n_sid, dtype, order, device):
    code, tdt = _DT[np.dtype(dtype)]
    if _order_code(order) == _lib.ORDER_F:
        base = torch.empty((n_sid, n_iid), dtype=tdt, device=device)
        return code, base, base.t()
    base = torch.empty((n_iid, n_sid), dtype=tdt, device=device)
    return code, base, base


def read(store, iid_sel=None, sid_sel=None, count_A1=False, dtype=np.float32, order="F", standardizer=None,
         stats=None, want_out=True):
    """Decode (and optionally standardize) a selection of the store on the GPU.

    Returns ``(val, stats)``: ``val`` a CUDA tensor ``[n_iid, n_sid]`` with the requested memory order
    (``None`` when ``want_out`` is False), ``stats`` a float64 CUDA tensor ``[n_sid, 2]`` (``None`` for a
    plain decode).  ``stats`` given => applied as trained statistics (UnitTrained / BetaTrained).
    """
    _lib.require_gpu()
    dev = store.device
    isel = iid_sel if isinstance(iid_sel, Selection) else Selection(iid_sel, store.iid_count, dev)
    ssel = sid_sel if isinstance(sid_sel, Selection) else Selection(sid_sel, store.sid_count, dev)
    mode, a, b = _mode_args(standardizer)
    with torch.cuda.device(dev):
        code, base, view = _alloc_out(isel.n, ssel.n, dtype, order, dev) if want_out else (_DT[np.dtype(dtype)][0], None, None)
        use_stats = 0
        d_stats = None
        if mode != _lib.STD_NONE:
            if stats is not None:
                d_stats = torch.as_tensor(np.asarray(stats, dtype=np.float64) if not torch.is_tensor(stats) else stats,
                                          dtype=torch.float64, device=dev).contiguous()
                assert tuple(d_stats.shape) == (ssel.n, 2), "stats must be [n_sid, 2]"
                use_stats = 1
            else:
                d_stats = torch.empty((ssel.n, 2), dtype=torch.float64, device=dev)
        check(lib.pstb_decode_standardize(
            store.tensor.data_ptr(), store.ld, store.iid_count, store.sid_count, isel.axis(), ssel.axis(), int(bool(count_A1)),
            mode, a, b, use_stats, d_stats.data_ptr() if d_stats is not None else None,
            base.data_ptr() if base is not None else None, code, _order_code(order), _stream()))
    return view, d_stats


def standardize(val, standardizer, stats=None, apply_in_place=True):
    """In-place Unit / Beta standardize of a CUDA tensor ``[n_iid, n_sid]`` (C- or F-contiguous). Returns stats."""
    _lib.require_gpu()
    mode, a, b = _mode_args(standardizer)
    n_iid, n_sid = val.shape
    if val.is_contiguous():
        order = _lib.ORDER_C
    elif val.t().is_contiguous():
        order = _lib.ORDER_F
    else:
        raise ValueError("val must be C- or F-contiguous")
    code = {torch.float32: _lib.F32, torch.float64: _lib.F64}[val.dtype]
    dev = val.device
    with torch.cuda.device(dev):
        use_stats = 0
        if stats is not None:
            d_stats = torch.as_tensor(stats, dtype=torch.float64, device=dev).contiguous().clone()
            use_stats = 1
        else:
            d_stats = torch.empty((n_sid, 2), dtype=torch.float64, device=dev)
        work = torch.empty(int(lib.pstb_standardize_work_bytes(n_sid)), dtype=torch.uint8, device=dev)
        check(lib.pstb_standardize(val.data_ptr(), code, order, n_iid, n_sid, mode, a, b, int(bool(apply_in_place)), use_stats,
                                   d_stats.data_ptr(), work.data_ptr(), _stream()))
    return d_stats


def pack(val, count_A1=False):
    """CUDA tensor ``[n_iid, n_sid]`` of {0,1,2,NaN / -127} -> PackedStore. Raises ValueError on other values."""
    _lib.require_gpu()
    n_iid, n_sid = val.shape
    if val.is_contiguous():
        order = _lib.ORDER_C
    elif val.t().is_contiguous():
        order = _lib.ORDER_F
    else:
        val, order = val.contiguous(), _lib.ORDER_C
    code = {torch.float32: _lib.F32, torch.float64: _lib.F64, torch.int8: _lib.I8}[val.dtype]
    dev = val.device
    with torch.cuda.device(dev):
        ld = int(lib.pstb_packed_ld(n_iid))
        out = torch.zeros((n_sid, max(ld, 16)), dtype=torch.uint8, device=dev)
        bad = torch.zeros(1, dtype=torch.int32, device=dev)
        check(lib.pstb_pack(val.data_ptr(), code, order, n_iid, n_sid, int(bool(count_A1)), out.data_ptr(), out.shape[1],
                            bad.data_ptr(), _stream()))
        if int(bad.item()) != 0:
            raise ValueError("Attempt to write illegal value to .bed: values must be 0, 1, 2 or missing (NaN / -127)")
    return PackedStore(out, n_iid, n_sid)


def snp_kernel(store, iid_sel=None, sid_sel=None, count_A1=False, standardizer=("unit",), stats=None, chunk=None,
               K=None, accumulate=False, mirror=True, low_term="default"):
    """``K = sum_j x_j x_j^T`` over the selected SNPs on tcgen05 tensor cores (float32 CUDA tensor [n, n]).

    Returns ``(K, stats)``.  ``K`` given + ``accumulate`` => the partial sum is added (streaming / sharding).
    ``low_term``: 'fp16' | 'fp8' | 'auto' | 'default' -- per call, see :func:`low_term_for`.
    """
    _lib.require_gpu()
    dev = store.device
    isel = iid_sel if isinstance(iid_sel, Selection) else Selection(iid_sel, store.iid_count, dev)
    ssel = sid_sel if isinstance(sid_sel, Selection) else Selection(sid_sel, store.sid_count, dev)
    mode, a, b = _mode_args(standardizer)
    n = isel.n
    with torch.cuda.device(dev):
        if K is None:
            K = torch.zeros((n, n), dtype=torch.float32, device=dev)
            accumulate = False
        assert K.dtype == torch.float32 and K.is_contiguous() and tuple(K.shape) == (n, n)
        use_stats = 0
        if stats is not None:
            d_stats = torch.as_tensor(stats, dtype=torch.float64, device=dev).contiguous()
            use_stats = 1
        else:
            d_stats = torch.empty((ssel.n, 2), dtype=torch.float64, device=dev)
        if chunk is None:
            chunk = default_kernel_chunk(n, ssel.n)
        wbytes = int(lib.pstb_kernel_workspace_bytes(n, chunk))
        work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
        check(lib.pstb_snp_kernel(store.tensor.data_ptr(), store.ld, store.iid_count, store.sid_count, isel.axis(), ssel.axis(),
                                  int(bool(count_A1)), mode, a, b, use_stats, d_stats.data_ptr(), K.data_ptr(),
                                  int(bool(accumulate)), int(bool(mirror)), work.data_ptr(), wbytes, chunk, _LOW_TERM[low_term], _stream()))
    return K, d_stats


def snp_cross_kernel(store_r, store_c, iid_r=None, iid_c=None, sid_r=None, sid_c=None, count_A1_r=False, count_A1_c=False,
                     standardizer=("unit",), stats=None, chunk=None, out=None, accumulate=False, low_term="default"):
    """Train x test kernel ``out[i, k] = sum_j x_ij y_kj`` (float32 CUDA tensor [n_r, n_c]) on the tensor cores.

    ``store_r`` / ``store_c`` are the row (train) and column (test) packed stores (may be the same store with different iid
    selections); both sides are standardized with the statistics of the row side (or with ``stats`` when given).
    Returns ``(out, stats)``.
    """
    _lib.require_gpu()
    dev = store_r.device
    assert store_c.device == dev, "both stores must live on the same GPU"
    ir = iid_r if isinstance(iid_r, Selection) else Selection(iid_r, store_r.iid_count, dev)
    ic = iid_c if isinstance(iid_c, Selection) else Selection(iid_c, store_c.iid_count, dev)
    sr = sid_r if isinstance(sid_r, Selection) else Selection(sid_r, store_r.sid_count, dev)
    sc = sid_c if isinstance(sid_c, Selection) else Selection(sid_c, store_c.sid_count, dev)
    if sr.n != sc.n:
        raise ValueError("both sides must select the same number of SNPs ({0} vs {1})".format(sr.n, sc.n))
    mode, a, b = _mode_args(standardizer)
    with torch.cuda.device(dev):
        if out is None:
            out = torch.zeros((ir.n, ic.n), dtype=torch.float32, device=dev)
            accumulate = False
        assert out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (ir.n, ic.n)
        use_stats = 0
        if stats is not None:
            d_stats = torch.as_tensor(stats, dtype=torch.float64, device=dev).contiguous()
            assert tuple(d_stats.shape) == (sr.n, 2), "stats must be [n_sid, 2]"
            use_stats = 1
        else:
            d_stats = torch.empty((sr.n, 2), dtype=torch.float64, device=dev)
        if chunk is None:
            chunk = default_kernel_chunk(ir.n + ic.n, sr.n)
        wbytes = int(lib.pstb_cross_kernel_workspace_bytes(ir.n, ic.n, chunk))
        work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
        check(lib.pstb_snp_cross_kernel(store_r.tensor.data_ptr(), store_r.ld, store_r.iid_count, store_r.sid_count, ir.axis(), sr.axis(),
                                        int(bool(count_A1_r)), store_c.tensor.data_ptr(), store_c.ld, store_c.iid_count, store_c.sid_count,
                                        ic.axis(), sc.axis(), int(bool(count_A1_c)), mode, a, b, use_stats, d_stats.data_ptr(),
                                        out.data_ptr(), int(bool(accumulate)), work.data_ptr(), wbytes, chunk, _LOW_TERM[low_term], _stream()))
    return out, d_stats


def kernel_tile_coords(n_iid, rank=0, world=1):
    """(I, J) block coordinates of the 256 x 256 lower-triangular tiles owned by ``rank`` of ``world`` (int32 [count, 2])."""
    count = int(lib.pstb_kernel_tile_count(n_iid, rank, world))
    ij = np.zeros((count, 2), dtype=np.int32)
    if count:
        check(lib.pstb_kernel_tile_coords(n_iid, rank, world, ij.ctypes.data))
    return ij


def snp_kernel_tiles(store, iid_sel=None, sid_sel=None, count_A1=False, standardizer=("unit",), stats=None, chunk=None,
                     rank=0, world=1, tiles=None, accumulate=False, low_term="default", defer_rank1=False):
    """K-tile sharded kinship: this rank's 256 x 256 tiles of ``K = X X^T`` over ALL selected SNPs (cfg5; no collective).

    Returns ``(tiles, coords, stats)``: float32 CUDA tensor [count, 256, 256], the (I, J) of each tile, per-SNP statistics.
    """
    _lib.require_gpu()
    dev = store.device
    isel = iid_sel if isinstance(iid_sel, Selection) else Selection(iid_sel, store.iid_count, dev)
    ssel = sid_sel if isinstance(sid_sel, Selection) else Selection(sid_sel, store.sid_count, dev)
    mode, a, b = _mode_args(standardizer)
    n = isel.n
    coords = kernel_tile_coords(n, rank, world)
    with torch.cuda.device(dev):
        if tiles is None:
            tiles = torch.zeros((len(coords), 256, 256), dtype=torch.float32, device=dev)
            accumulate = False
        assert tiles.dtype == torch.float32 and tiles.is_contiguous() and tuple(tiles.shape) == (len(coords), 256, 256)
        use_stats = 0
        if stats is not None:
            d_stats = torch.as_tensor(stats, dtype=torch.float64, device=dev).contiguous()
            use_stats = 1
        else:
            d_stats = torch.empty((ssel.n, 2), dtype=torch.float64, device=dev)
        if chunk is None:
            chunk = default_kernel_chunk(n, ssel.n)
        wbytes = int(lib.pstb_kernel_workspace_bytes(n, chunk))
        work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
        check(lib.pstb_snp_kernel_tiles(store.tensor.data_ptr(), store.ld, store.iid_count, store.sid_count, isel.axis(), ssel.axis(),
                                        int(bool(count_A1)), mode, a, b, use_stats, d_stats.data_ptr(), tiles.data_ptr(), rank, world,
                                        int(bool(accumulate)) | (2 if defer_rank1 else 0), work.data_ptr(), wbytes, chunk, _LOW_TERM[low_term], _stream()))
        if defer_rank1:
            # the rank-one vector v of the exact-dosage path stays in the workspace: hand a copy to the caller (multi-GPU: all-reduced with the tiles)
            u = workspace_rank1(work, n, chunk).clone() if (n and ssel.n) else torch.zeros(n, dtype=torch.float64, device=dev)
            return tiles, coords, d_stats, u
    return tiles, coords, d_stats


def workspace_rank1(work, n_iid, chunk):
    """float64 view [n_iid] of the rank-one vector inside a kernel workspace (``pstb_kernel_workspace_rank1``)."""
    ptr = lib.pstb_kernel_workspace_rank1(work.data_ptr(), int(n_iid), int(chunk))
    if not ptr:
        raise _lib.PstB200Error("no rank-one vector in this workspace")
    off = int(ptr) - int(work.data_ptr())
    return work[off:off + 8 * int(n_iid)].view(torch.float64)


def set_syrk_low_term(mode):
    """'fp16' | 'fp8' | 'auto': the process-wide DEFAULT of the kernel entry points' per-call ``low_term`` argument
    (``pstb_set_syrk_low_term``).  Returns the previous default's name."""
    prev = int(lib.pstb_set_syrk_low_term(_LOW_TERM[mode]))
    return _LOW_TERM_NAME[prev]


def get_syrk_low_term():
    return _LOW_TERM_NAME[int(lib.pstb_set_syrk_low_term(-1))]        # an invalid mode only reports


def low_term_for(total_sid_count, n_iid, standardizer=("unit",)):
    """The ``low_term`` a kernel call should pass when the SNPs of ONE kernel are spread over several calls (SNP shards,
    DistributedBed pieces, file slices): 'auto' inside the library looks at one call's SNP count, this applies the same rule to
    the count of the whole kernel.  An explicit process-wide 'fp16' / 'fp8' default is respected.  No global state is touched,
    so concurrent callers with different shapes cannot race (the round-1 context manager switched a process-wide mode)."""
    default = get_syrk_low_term()
    if default != "auto":
        return default
    need = n_iid * (4 if standardizer is not None and standardizer[0] == "beta" else 1)
    return "fp8" if (total_sid_count >= need and total_sid_count >= 256) else "fp16"


def kernel_from_tiles(tiles, n_iid, rank=0, world=1, K=None):
    """Compact tiles ``[count, 256, 256]`` (:func:`snp_kernel_tiles` layout) -> full symmetric float32 ``K`` [n, n]."""
    _lib.require_gpu()
    dev = tiles.device
    with torch.cuda.device(dev):
        if K is None:
            K = torch.zeros((n_iid, n_iid), dtype=torch.float32, device=dev)
        assert K.dtype == torch.float32 and K.is_contiguous() and tuple(K.shape) == (n_iid, n_iid) and tiles.is_contiguous()
        check(lib.pstb_kernel_from_tiles(tiles.data_ptr(), n_iid, rank, world, K.data_ptr(), _stream()))
    return K


def float_kernel(val, chunk=None, K=None, accumulate=False, mirror=True):
    """``K = V V^T`` of a CUDA float tensor [n_iid, n_sid] (C- or F-contiguous) on the tensor cores (float32 result)."""
    _lib.require_gpu()
    n, m = val.shape
    if val.is_contiguous():
        order = _lib.ORDER_C
    elif val.t().is_contiguous():
        order = _lib.ORDER_F
    else:
        val, order = val.contiguous(), _lib.ORDER_C
    code = {torch.float32: _lib.F32, torch.float64: _lib.F64}[val.dtype]
    dev = val.device
    with torch.cuda.device(dev):
        if K is None:
            K = torch.zeros((n, n), dtype=torch.float32, device=dev)
            accumulate = False
        if chunk is None:
            chunk = default_kernel_chunk(n, m)
        wbytes = int(lib.pstb_kernel_workspace_bytes(n, chunk))
        work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
        check(lib.pstb_float_kernel(val.data_ptr(), code, order, n, m, K.data_ptr(), int(bool(accumulate)), int(bool(mirror)),
                                    work.data_ptr(), wbytes, chunk, _stream()))
    return K


def snp_kernel_f64(store, iid_sel=None, sid_sel=None, count_A1=False, standardizer=("unit",), stats=None, chunk=None,
                   K=None, accumulate=False, mirror=True):
    """``K = sum_j x_j x_j^T`` in float64 arithmetic (``pstb_snp_kernel_f64``: fused decode + standardize into a float64 panel, fp64-FMA
    SYRK): what a ``dtype=float64`` kernel of the reference means (snpdata.py:203-206 is a DGEMM).  float64 CUDA tensor [n, n]."""
    _lib.require_gpu()
    dev = store.device
    isel = iid_sel if isinstance(iid_sel, Selection) else Selection(iid_sel, store.iid_count, dev)
    ssel = sid_sel if isinstance(sid_sel, Selection) else Selection(sid_sel, store.sid_count, dev)
    mode, a, b = _mode_args(standardizer)
    n = isel.n
    with torch.cuda.device(dev):
        if K is None:
            K = torch.zeros((n, n), dtype=torch.float64, device=dev)
            accumulate = False
        assert K.dtype == torch.float64 and K.is_contiguous() and tuple(K.shape) == (n, n)
        use_stats = 0
        if stats is not None:
            d_stats = torch.as_tensor(stats, dtype=torch.float64, device=dev).contiguous()
            use_stats = 1
        else:
            d_stats = torch.empty((ssel.n, 2), dtype=torch.float64, device=dev)
        if chunk is None:
            chunk = max(64, min(4096, (256 << 20) // max(1, 8 * n) // 64 * 64))        # a float64 panel of ~256 MB
        wbytes = int(lib.pstb_kernel_f64_workspace_bytes(n, chunk))
        work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
        check(lib.pstb_snp_kernel_f64(store.tensor.data_ptr(), store.ld, store.iid_count, store.sid_count, isel.axis(), ssel.axis(),
                                      int(bool(count_A1)), mode, a, b, use_stats, d_stats.data_ptr(), K.data_ptr(),
                                      int(bool(accumulate)), int(bool(mirror)), work.data_ptr(), wbytes, chunk, _stream()))
    return K, d_stats


def float_kernel_f64(val, K=None, accumulate=False, mirror=True):
    """``K = V V^T`` of a float64 CUDA tensor [n_iid, n_sid] (C- or F-contiguous) in float64 arithmetic (``pstb_float_kernel_f64``)."""
    _lib.require_gpu()
    assert val.dtype == torch.float64, "the float64 kernel path takes float64 values"
    n, m = val.shape
    if val.is_contiguous():
        order = _lib.ORDER_C
    elif val.t().is_contiguous():
        order = _lib.ORDER_F
    else:
        val, order = val.contiguous(), _lib.ORDER_C
    dev = val.device
    with torch.cuda.device(dev):
        if K is None:
            K = torch.zeros((n, n), dtype=torch.float64, device=dev)
            accumulate = False
        check(lib.pstb_float_kernel_f64(val.data_ptr(), order, n, m, K.data_ptr(), int(bool(accumulate)), int(bool(mirror)), _stream()))
    return K


def default_kernel_chunk(n_iid, n_sid):
    """SNPs per operand-plane chunk, a multiple of 64.

    The SYRK walks the lower triangle in 2048 x 2048 super-blocks; a chunk is sized so that the operand rows of one
    super-block (2048 A rows x 1 plane + 2048 B rows x 2 planes, fp16) stay L2-resident (~50 MB of the 126 MB):
    measured on cfg3, 4096-SNP chunks give 1048 TFLOP/s, 8192: 1015, 16384: 980, 32768: 946, 2048: 994.
    """
    n_pad = max(256, ((n_iid + 255) // 256) * 256)
    rows = min(n_pad, 2048)
    chunk = (50_000_000 // (6 * rows)) // 64 * 64
    chunk = max(1024, min(16384, chunk))
    return int(min(chunk, max(64, (n_sid + 63) // 64 * 64)))


def convert_kernel(K32, dtype=np.float64, scale=1.0):
    """float32 K -> float32 / float64 CUDA tensor with an optional scalar factor (DiagKtoN)."""
    n = K32.shape[0]
    code, tdt = _DT[np.dtype(dtype)]
    out = torch.empty((n, n), dtype=tdt, device=K32.device)
    with torch.cuda.device(K32.device):
        check(lib.pstb_convert_kernel(K32.data_ptr(), n, out.data_ptr(), code, float(scale), _stream()))
    return out
