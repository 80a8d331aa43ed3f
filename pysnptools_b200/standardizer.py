"""Standardizers: the reference's ``pysnptools.standardizer`` surface for the hot path, backed by CUDA.

Mirrors (same names / arguments / return conventions):
``Unit`` (standardizer/unit.py:28-51), ``Beta`` (beta.py:33-50), ``UnitTrained`` (unittrained.py:47-70),
``BetaTrained`` (betatrained.py:47-63), ``Identity`` (identity.py), ``DiagKtoN`` (diag_K_to_N.py:54-95).
The dispatch the reference does in ``Standardizer._standardize_unit_and_beta`` (standardizer.py:89-133,
to the Rust ``standardize_f32/f64``) goes to ``pstb_standardize_host`` (host arrays) or ``pstb_standardize``
(CUDA tensors).  ``force_python_only=True`` raises: this package has no CPU path.
"""
import warnings

import numpy as np

from . import _lib


def _no_python_path(force_python_only):
    if force_python_only:
        raise NotImplementedError("pysnptools_b200 has no pure-Python / CPU path (force_python_only=True); "
                                  "use the reference package for that")


def _is_tensor(x):
    return type(x).__module__.startswith("torch")


def _standardize_unit_and_beta(val, is_beta, a, b, apply_in_place, use_stats, stats, num_threads=None, force_python_only=False):
    """In-place standardize of ``val`` [iid, sid]; returns stats [sid, 2] in ``val``'s dtype (standardizer.py:89-133)."""
    _no_python_path(force_python_only)
    mode = _lib.STD_BETA if is_beta else _lib.STD_UNIT
    if _is_tensor(val):
        from . import device
        st = device.standardize(val, ("beta", a, b) if is_beta else ("unit",), stats=stats if use_stats else None,
                                apply_in_place=apply_in_place)
        return st.to(val.dtype)
    assert val.dtype in (np.float32, np.float64), "snps must be a float in order to standardize in place."
    assert val.flags["C_CONTIGUOUS"] or val.flags["F_CONTIGUOUS"], "Expect snps to be order 'C' or order 'F'"
    if apply_in_place and not val.flags.writeable:
        # the C ABI writes through the raw pointer: a read-only mapping (SnpMemMap opens its file with mode='r') would take the
        # process down, a writeable=False ndarray would be mutated silently; NumPy's in-place arithmetic raises exactly this
        raise ValueError("assignment destination is read-only")
    _lib.require_gpu()
    n_iid, n_sid = val.shape
    st64 = np.empty((n_sid, 2), dtype=np.float64)
    if use_stats:
        st64[...] = np.asarray(stats, dtype=np.float64)
        assert st64.shape == (n_sid, 2), "stats must have size [sid_count,2]"
    order = _lib.ORDER_C if val.flags["C_CONTIGUOUS"] else _lib.ORDER_F      # both set only for degenerate shapes: same memory
    code = _lib.F64 if val.dtype == np.float64 else _lib.F32
    _lib.check(_lib.lib.pstb_standardize_host(val.ctypes.data, code, order, n_iid, n_sid, mode, float(a), float(b),
                                              int(bool(apply_in_place)), int(bool(use_stats)), st64.ctypes.data))
    return st64.astype(val.dtype, copy=False)


class Standardizer(object):
    """Base class (standardizer/standardizer.py:59-86)."""

    def standardize(self, snps, block_size=None, return_trained=False, force_python_only=False, num_threads=None):
        raise NotImplementedError("subclass {0} needs to implement method '.standardize'".format(self.__class__.__name__))

    @property
    def is_constant(self):
        return False

    def _merge_trained(self, trained_list):
        raise Exception("Not defined")

    # what the fused GPU kernels need: None (identity) | ("unit",) | ("beta", a, b)
    def _device_spec(self):
        return None                                     # no fused GPU path: callers read first, then call .standardize

    def _trained_stats_for(self, sid):
        return None

    def _make_trained(self, sid, stats):
        return self


def _val_of(snps, who):
    if hasattr(snps, "val"):
        return snps.val
    warnings.warn("standardizing an ndarray instead of a SnpData is deprecated", DeprecationWarning)
    return snps


def _warn_block_size(block_size):
    if block_size is not None:
        warnings.warn("block_size is deprecated (and not needed, since standardization is in-place", DeprecationWarning)


class Unit(Standardizer):
    """Mean 0, sd 1 per SNP; missing -> 0; SNC SNPs -> 0 (standardizer/unit.py)."""

    def __repr__(self):
        return "{0}()".format(self.__class__.__name__)

    def standardize(self, snps, block_size=None, return_trained=False, force_python_only=False, num_threads=None):
        _warn_block_size(block_size)
        val = _val_of(snps, self)
        stats = _standardize_unit_and_beta(val, False, np.nan, np.nan, True, False, None, num_threads, force_python_only)
        if return_trained:
            assert hasattr(snps, "val"), "return_trained=True requires that snps be a SnpData"
            return snps, UnitTrained(snps.sid, stats)
        return snps

    def _merge_trained(self, trained_list):
        sid = np.concatenate([t.sid for t in trained_list])
        stats = np.concatenate([np.asarray(t.stats) for t in trained_list])
        return UnitTrained(sid, stats)

    def _device_spec(self):
        return ("unit",)

    def _make_trained(self, sid, stats):
        return UnitTrained(sid, stats)


class Beta(Standardizer):
    """Centre and weight by BetaPDF(maf; a, b); missing -> 0 (standardizer/beta.py)."""

    def __init__(self, a, b):
        super(Beta, self).__init__()
        self.a, self.b = a, b

    def __repr__(self):
        return "{0}(a={1},b={2})".format(self.__class__.__name__, self.a, self.b)

    def standardize(self, snps, block_size=None, return_trained=False, force_python_only=False, num_threads=None):
        _warn_block_size(block_size)
        val = _val_of(snps, self)
        stats = _standardize_unit_and_beta(val, True, self.a, self.b, True, False, None, num_threads, force_python_only)
        if return_trained:
            assert hasattr(snps, "val"), "return_trained=True requires that snps be a SnpData"
            return snps, BetaTrained(self.a, self.b, snps.sid, stats)
        return snps

    def _merge_trained(self, trained_list):
        sid = np.concatenate([t.sid for t in trained_list])
        stats = np.concatenate([np.asarray(t.stats) for t in trained_list])
        return BetaTrained(self.a, self.b, sid, stats)

    def _device_spec(self):
        return ("beta", self.a, self.b)

    def _make_trained(self, sid, stats):
        return BetaTrained(self.a, self.b, sid, stats)


class _Trained(Standardizer):
    def __init__(self, sid, stats):
        super(_Trained, self).__init__()
        self.sid = sid
        self.stats = stats
        self.sid_to_index = None

    @property
    def is_constant(self):
        return True

    def _stats_for(self, snps):
        if hasattr(snps, "sid"):
            return self._trained_stats_for(snps.sid)
        return self.stats

    def _trained_stats_for(self, sid):
        if len(self.sid) == len(sid) and np.array_equal(self.sid, sid):
            return self.stats
        if self.sid_to_index is None:
            self.sid_to_index = {s: i for i, s in enumerate(self.sid)}
        return np.array([np.asarray(self.stats)[self.sid_to_index[s]] for s in sid])

    def _make_trained(self, sid, stats):
        return self


class UnitTrained(_Trained):
    """Unit with pre-computed [mean, sd] per SNP (standardizer/unittrained.py:47-70)."""

    def __repr__(self):
        return "{0}(stats={1},sid={2})".format(self.__class__.__name__, self.stats, self.sid)

    def standardize(self, snps, block_size=None, return_trained=False, force_python_only=False, num_threads=None):
        _warn_block_size(block_size)
        val = _val_of(snps, self)
        _standardize_unit_and_beta(val, False, np.nan, np.nan, True, True, self._stats_for(snps), num_threads, force_python_only)
        return (snps, self) if return_trained else snps

    def _device_spec(self):
        return ("unit",)


class BetaTrained(_Trained):
    """Beta with pre-computed [mean, sd] per SNP (standardizer/betatrained.py:47-63)."""

    def __init__(self, a, b, sid, stats):
        super(BetaTrained, self).__init__(sid, stats)
        self.a, self.b = a, b

    def __repr__(self):
        return "{0}(a={1},b={2},stats={3},sid={4})".format(self.__class__.__name__, self.a, self.b, self.stats, self.sid)

    def standardize(self, snps, block_size=None, return_trained=False, force_python_only=False, num_threads=None):
        _warn_block_size(block_size)
        val = _val_of(snps, self)
        _standardize_unit_and_beta(val, True, self.a, self.b, True, True, self._stats_for(snps), num_threads, force_python_only)
        return (snps, self) if return_trained else snps

    def _device_spec(self):
        return ("beta", self.a, self.b)


class Identity(Standardizer):
    """Leaves the values alone (standardizer/identity.py)."""

    def __repr__(self):
        return "{0}()".format(self.__class__.__name__)

    @property
    def is_constant(self):
        return True

    def standardize(self, snps, block_size=None, return_trained=False, force_python_only=False, num_threads=None):
        _warn_block_size(block_size)
        return (snps, self) if return_trained else snps

    def _merge_trained(self, trained_list):
        return self

    def _device_spec(self):
        return None


class DiagKtoNTrained(Standardizer):
    def __init__(self, factor):
        super(DiagKtoNTrained, self).__init__()
        self.factor = factor

    @property
    def is_constant(self):
        return True

    def standardize(self, input, block_size=None, return_trained=False, force_python_only=False, num_threads=None):
        val = input.val if hasattr(input, "val") else input
        is_kernel = hasattr(input, "iid0") or getattr(input, "_is_kernel", False)
        f = self.factor if is_kernel else np.sqrt(self.factor)
        if abs(self.factor - 1.0) > 1e-15:
            val *= f
        return (input, self) if return_trained else input


class DiagKtoN(Standardizer):
    """Scale so that the kernel's diagonal sums to iid_count (standardizer/diag_K_to_N.py:54-95). O(N^2) scalar work on the host."""

    def __init__(self, deprecated_iid_count=None):
        super(DiagKtoN, self).__init__()
        if deprecated_iid_count is not None:
            warnings.warn("'iid_count' is deprecated (and not needed, since can get iid_count from SNPs val's first dimension", DeprecationWarning)

    def __repr__(self):
        return "{0}()".format(self.__class__.__name__)

    def standardize(self, input, block_size=None, return_trained=False, force_python_only=False, num_threads=None):
        _warn_block_size(block_size)
        if getattr(input, "_is_kernel", False):
            val = input.val
            diag_sum = float(val.diagonal().sum())
            factor = float(input.iid_count) / diag_sum
            if abs(factor - 1.0) > 1e-15:
                val *= factor
        else:
            val = _val_of(input, self)
            if _is_tensor(val):
                squared_sum = float((val.double() ** 2).sum().item())
            else:
                vec = val.reshape(-1, order="A")
                squared_sum = float(vec.dot(vec))
            factor = float(val.shape[0]) / squared_sum
            if abs(factor - 1.0) > 1e-15:
                val *= np.sqrt(factor)
        return (input, DiagKtoNTrained(factor)) if return_trained else input
