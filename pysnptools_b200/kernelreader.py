"""SnpKernel / KernelData: the reference's ``pysnptools.kernelreader`` surface for the hot path.

Mirrors ``SnpKernel`` (kernelreader/snpkernel.py:43-132), ``KernelReader.read`` (kernelreader.py:245-302) and
``KernelData`` (kerneldata.py:71-159).  The kernel itself is computed by ``SnpReader._read_kernel`` on the GPU.
"""
import numpy as np

from .snpreader import _compose, _resolve_indexer
from .standardizer import DiagKtoN, _is_tensor


class KernelReader(object):
    """Accessor surface shared by KernelData and SnpKernel (kernelreader/kernelreader.py:60-243); subclasses give row / col."""

    @property
    def iid0(self):
        return self.row

    @property
    def iid1(self):
        return self.col

    @property
    def iid(self):
        assert self.row is self.col or np.array_equal(self.row, self.col), "When 'iid' is used, iid0 must be the same as iid1"
        return self.row

    @property
    def iid_count(self):
        return len(self.iid)

    @property
    def iid0_count(self):
        return len(self.row)

    @property
    def iid1_count(self):
        return len(self.col)

    row_count = iid0_count
    col_count = iid1_count

    @property
    def shape(self):
        return (self.iid0_count, self.iid1_count)

    @property
    def row_property(self):
        return np.empty((self.iid0_count, 0))

    @property
    def col_property(self):
        return np.empty((self.iid1_count, 0))

    @property
    def val_shape(self):
        return None

    @staticmethod
    def _to_index(labels, wanted):
        lookup = {tuple(x): i for i, x in enumerate(labels)}
        return np.array([lookup[tuple(x)] for x in wanted], dtype=np.int64)

    def iid0_to_index(self, list):
        return self._to_index(self.row, list)

    def iid1_to_index(self, list):
        return self._to_index(self.col, list)

    def iid_to_index(self, list):
        return self._to_index(self.iid, list)

    row_to_index = iid0_to_index
    col_to_index = iid1_to_index

    def copyinputs(self, copier):
        pass


class KernelData(KernelReader):
    """In-memory kernel: ``val`` [iid0_count, iid1_count] + iids."""
    _is_kernel = True

    def __init__(self, iid=None, iid0=None, iid1=None, val=None, name=None, parent_string=None, xp=None):
        assert (iid is None) != (iid0 is None and iid1 is None), "Either 'iid' or both 'iid0' 'iid1' must be provided."
        assert val is not None, "'val' must not be None"
        if iid is not None:
            self._row = self._col = np.array(iid, dtype=str).reshape(-1, 2)
        else:
            self._row = np.array(iid0, dtype=str).reshape(-1, 2)
            self._col = np.array(iid1, dtype=str).reshape(-1, 2)
        assert tuple(val.shape) == (len(self._row), len(self._col)), "val shape must match the iid counts"
        self._val = val
        self._name = name or parent_string or ""
        self._std_string_list = []

    def __repr__(self):
        parts = ([self._name] if self._name else []) + self._std_string_list
        return "KernelData({0})".format(",".join(parts))

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

    def allclose(self, value, equal_nan=True):
        """Same iids and close values (kerneldata.py:113-134)."""
        def arr(v):
            return v.cpu().numpy() if _is_tensor(v) else np.asarray(v)
        return (np.array_equal(self.iid0, value.iid0) and np.array_equal(self.iid1, value.iid1)
                and bool(np.allclose(arr(self.val), arr(value.val), equal_nan=equal_nan)))

    def read(self, order="F", dtype=np.float64, force_python_only=False, view_ok=False, num_threads=None):
        val = self._val
        if _is_tensor(val):
            val = val.cpu().numpy()
        dtype = np.dtype(dtype)
        ok = order == "A" or val.flags["F_CONTIGUOUS" if order == "F" else "C_CONTIGUOUS"]
        if not (view_ok and ok and val.dtype == dtype):
            val = np.array(val, dtype=dtype, order=order)
        return KernelData(iid0=self._row, iid1=self._col, val=val, name=str(self))

    def __getitem__(self, iid_indexer_and_snp_indexer):
        indexer = iid_indexer_and_snp_indexer
        r, c = indexer if isinstance(indexer, tuple) else (indexer, indexer)
        ri, ci = _resolve_indexer(r, self.iid0_count), _resolve_indexer(c, self.iid1_count)
        val = self._val
        if ri is not None:
            val = val[ri]
        if ci is not None:
            val = val[:, ci]
        return KernelData(iid0=self._row if ri is None else self._row[ri], iid1=self._col if ci is None else self._col[ci], val=val)

    def standardize(self, standardizer=DiagKtoN(), return_trained=False, force_python_only=False, num_threads=None):
        """In place; the diagonal sums to iid_count afterwards (kerneldata.py:136-159)."""
        self._std_string_list.append(str(standardizer))
        return standardizer.standardize(self, return_trained=return_trained, force_python_only=force_python_only, num_threads=num_threads)


class SnpKernel(KernelReader):
    """Lazy ``K = X X^T`` of a standardized :class:`SnpReader` (kernelreader/snpkernel.py)."""

    def __init__(self, snpreader, standardizer=None, block_size=None, test=None):
        """``test=`` (keyword, an extension over snpkernel.py:43-50): a second SnpReader over the same SNPs; the kernel is then
        the train x test matrix FaST-LMM predicts with -- ``iid0`` = ``snpreader.iid``, ``iid1`` = ``test.iid``,
        ``val = X_train X_test^T`` with BOTH sides standardized by the statistics learned on ``snpreader`` (the *Trained
        standardizers, unittrained.py:47-70 / betatrained.py:47-63)."""
        assert standardizer is not None, "'standardizer' must be provided"
        self.snpreader = snpreader
        self.standardizer = standardizer
        self.block_size = block_size
        self.test = test
        self._index = None          # subset of the kernel's iids applied AFTER the kernel is computed

    def __repr__(self):
        s = "SnpKernel({0},standardizer={1}".format(self.snpreader, self.standardizer)
        if self.block_size is not None:
            s += ",block_size={0}".format(self.block_size)
        if self.test is not None:
            s += ",test={0}".format(self.test)
        return s + ")"

    @property
    def row(self):
        iid = self.snpreader.iid
        return iid if self._index is None else iid[self._index]

    @property
    def col(self):
        return self.row if self.test is None else self.test.iid

    # the SNP side of the underlying reader (snpkernel.py:134-150)
    @property
    def sid(self):
        return self.snpreader.sid

    @property
    def sid_count(self):
        return self.snpreader.sid_count

    @property
    def pos(self):
        return self.snpreader.pos

    def copyinputs(self, copier):
        copier.input(self.snpreader)
        copier.input(self.standardizer)

    def read_snps(self, order="F", dtype=np.float64, force_python_only=False, view_ok=False, num_threads=None):
        """The standardized SNP values behind the kernel as a SnpData (snpkernel.py:152-187): one fused GPU pass."""
        data = self.snpreader.read(order=order, dtype=dtype, force_python_only=force_python_only, view_ok=view_ok,
                                   num_threads=num_threads, standardizer=self.standardizer)
        return data if self._index is None else data[self._index, :].read(order=order, dtype=dtype, view_ok=True)

    def __getitem__(self, iid_indexer_and_snp_indexer):
        indexer = iid_indexer_and_snp_indexer
        r, c = indexer if isinstance(indexer, tuple) else (indexer, indexer)
        if self.test is not None:
            # rectangular: the column subset is a plain subset of the test reader; a row subset changes the training set only
            # for a non-constant standardizer, so it is applied after the read
            ri, ci = _resolve_indexer(r, self.iid0_count), _resolve_indexer(c, self.iid1_count)
            out = SnpKernel(self.snpreader, self.standardizer, block_size=self.block_size, test=self.test if ci is None else self.test[ci, :])
            out._index = _compose(self._index, ri) if ri is not None else self._index
            return out
        ri, ci = _resolve_indexer(r, self.iid_count), _resolve_indexer(c, self.iid_count)
        same = (ri is None and ci is None) or (ri is not None and ci is not None and np.array_equal(ri, ci))
        assert same, "SnpKernel supports the same indexer on both axes"
        if ri is None:
            return self
        if self.standardizer.is_constant and self._index is None:
            # constant standardizer: push the subset into the reader (snpkernel.py:90-101)
            return SnpKernel(self.snpreader[ri, :], self.standardizer, block_size=self.block_size)
        out = SnpKernel(self.snpreader, self.standardizer, block_size=self.block_size)
        out._index = _compose(self._index, ri)
        return out

    def _read_cross(self, order, dtype, force_python_only, return_trained):
        """Train x test on the GPU: pstb_snp_cross_kernel over the two packed stores."""
        from . import device
        from .standardizer import Standardizer, _no_python_path
        from .snpreader import _kernel_chunk
        _no_python_path(force_python_only)
        std = self.standardizer
        if not isinstance(std, Standardizer) or std._device_spec() is None:
            raise NotImplementedError("train x test kernels on the GPU support Unit, Beta and their trained forms")
        train, test = self.snpreader, self.test
        if not np.array_equal(train.sid, test.sid):
            test = test[:, test.sid_to_index(train.sid)]                        # pair the SNPs by name (a KeyError names a missing one)
        root_r, ii_r, si_r = train._root_and_indices()
        root_c, ii_c, si_c = test._root_and_indices()
        store_r, sel_r = root_r._store_for(si_r)
        store_c, sel_c = root_c._store_for(si_c)
        stats = std._trained_stats_for(train.sid)
        out, d_stats = device.snp_cross_kernel(store_r, store_c, ii_r, ii_c, sel_r, sel_c, count_A1_r=root_r.count_A1,
                                               count_A1_c=root_c.count_A1, standardizer=std._device_spec(), stats=stats,
                                               chunk=_kernel_chunk(self.block_size, train.iid_count + test.iid_count, train.sid_count))
        val = out.cpu().numpy().astype(dtype, copy=False)
        if self._index is not None:
            val = val[self._index]
        val = np.asarray(val, order="F" if order == "F" else "C")
        if return_trained:
            return val, std._make_trained(train.sid, d_stats.cpu().numpy().astype(dtype if dtype in (np.float32, np.float64) else np.float64))
        return val

    def _read(self, order, dtype, force_python_only, view_ok, num_threads, return_trained=False):
        if self.test is not None:
            return self._read_cross(order, np.dtype(dtype), force_python_only, return_trained)
        res = self.snpreader._read_kernel(self.standardizer, self.block_size, order, dtype, force_python_only, view_ok,
                                          return_trained=return_trained, num_threads=num_threads)
        val, trained = res if return_trained else (res, None)
        if self._index is not None:
            # standardize on all iids, then slice (kernelreader/test.py:235-247)
            val = np.array(val[self._index][:, self._index], order="F" if order == "F" else "C")
        return (val, trained) if return_trained else val

    def read(self, order="F", dtype=np.float64, force_python_only=False, view_ok=False, num_threads=None):
        val = self._read(order, np.dtype(dtype), force_python_only, view_ok, num_threads)
        if self.test is not None:
            return KernelData(iid0=self.row, iid1=self.col, val=val, name=str(self))
        return KernelData(iid=self.row, val=val, name=str(self))

    def _read_with_standardizing(self, to_kerneldata, kernel_standardizer=DiagKtoN(), return_trained=False, num_threads=None):
        """FaST-LMM's entry (snpkernel.py:104-132): kernel + trained SNP standardizer + trained kernel standardizer."""
        assert to_kerneldata, "only the KernelData form is on the GPU path"
        val, snp_trained = self._read("A", np.dtype(np.float64), False, False, num_threads, return_trained=True)
        if self.test is not None:
            # a train x test kernel has no diagonal to learn from: it takes the factor learned on the train kernel
            assert kernel_standardizer.is_constant, "a train x test kernel needs a trained kernel standardizer (DiagKtoNTrained) or Identity"
            kernel = KernelData(iid0=self.row, iid1=self.col, val=val, name=str(self))
        else:
            kernel = KernelData(iid=self.row, val=val, name=str(self))
        kernel, kernel_trained = kernel.standardize(kernel_standardizer, return_trained=True, num_threads=num_threads)
        return (kernel, snp_trained, kernel_trained) if return_trained else kernel
