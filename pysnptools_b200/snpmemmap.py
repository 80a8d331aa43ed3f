"""SnpMemMap: SNP values in a memory-mapped file -- the host-side staging format either side of the GPU path.

Mirrors ``pysnptools/snpreader/snpmemmap.py`` (``SnpMemMap``: 44-202) on top of the file format of
``pysnptools/pstreader/pstmemmap.py`` (``_empty_inner`` 156-183, ``_run_once_inner`` 192-223): a sequence of ``np.save`` records
(magic 22891, "pstmemmap", version 2, row, col, row_property, col_property, dtype, order, val_shape) followed by the raw matrix,
which ``np.memmap`` maps at that offset.  Files written here open in the reference and the other way round
(``tests/test_host_logic.py::test_snpmemmap_interop_with_reference``).

What runs where: ``SnpMemMap.write(filename, bed, standardizer)`` streams SNP blocks through the fused decode + standardize
kernel and lands them in the mapped file, so a standardized matrix larger than host RAM can be staged on disk;
``read_kernel`` streams column blocks of the mapped matrix through ``pstb_standardize`` / ``pstb_float_kernel``.  Subsetting a
mapped matrix is done by the host (it is file I/O: only the pages of the selected rows / columns are touched).
"""
import os
import shutil

import numpy as np

from .snpreader import SnpData, SnpReader, _symmetric_to_host
from .standardizer import Identity, Standardizer, _is_tensor, _no_python_path

_MAGIC = 22891


class SnpMemMap(SnpData):
    def __init__(self, filename):
        SnpReader.__init__(self)
        self._filename = str(filename)
        self._ran_once = False
        self._std_string_list = []
        self._name = "np.memmap('{0}')".format(self._filename)

    def __repr__(self):
        return "{0}('{1}')".format(self.__class__.__name__, self._filename)

    def __getstate__(self):
        return self._filename

    def __setstate__(self, state):
        self.__init__(state)

    # ---- lazily opened file ---------------------------------------------------------------------------------------
    def _run_once(self):
        if self._ran_once:
            return
        with open(self._filename, "rb") as fp:
            first = np.load(fp, allow_pickle=True)
            if len(first) == 1 and first[0] == _MAGIC:
                fmt = np.load(fp, allow_pickle=True)[0]
                version = np.load(fp, allow_pickle=True)[0]
                assert fmt == "pstmemmap", "Expect format of 'pstmemmap'"
                assert version == 2, "Expect version of 2"
                row = np.load(fp, allow_pickle=True)
            else:                                               # version 1 starts with the row array itself
                row = first
            col = np.load(fp, allow_pickle=True)
            np.load(fp, allow_pickle=True)                       # row_property: SnpReaders carry none
            col_property = np.load(fp, allow_pickle=True)
            dtype = np.dtype(np.load(fp, allow_pickle=True)[0])
            order = str(np.load(fp, allow_pickle=True)[0])
            val_shape = np.load(fp, allow_pickle=True)[0] if len(first) == 1 and first[0] == _MAGIC else None
            offset = fp.tell()
        assert val_shape is None, "SnpMemMap holds 2-D values"
        val = np.memmap(self._filename, offset=offset, dtype=dtype, mode="r", order=order, shape=(len(row), len(col)))
        self._attach(row, col, col_property, val, offset, dtype, order)

    def _attach(self, row, col, col_property, val, offset, dtype, order):
        self._row = np.array(row, dtype=str).reshape(-1, 2)
        self._col = np.array(col, dtype=str).reshape(-1)
        self._col_property = np.array(col_property, dtype=np.float64).reshape(-1, 3)
        self._val, self._offset, self._dtype, self._order = val, offset, dtype, order
        self._ran_once = True

    @property
    def row(self):
        self._run_once()
        return self._row

    @property
    def col(self):
        self._run_once()
        return self._col

    @property
    def col_property(self):
        self._run_once()
        return self._col_property

    @property
    def val(self):
        """The memory-mapped matrix.  It can be written through (``snp_mem_map.val[:, :] = ...``) but not replaced."""
        self._run_once()
        return self._val

    @val.setter
    def val(self, new_value):
        self._run_once()
        if self._val is new_value:
            return
        raise Exception("SnpMemMap val's cannot be set to a different array")

    @property
    def offset(self):
        """Byte position in the file where the matrix starts."""
        self._run_once()
        return self._offset

    @property
    def filename(self):
        return self._filename

    def copyinputs(self, copier):
        copier.input(self._filename)

    # ---- creating files ----------------------------------------------------------------------------------------------
    @staticmethod
    def empty(iid, sid, filename, pos=None, order="F", dtype=np.float64):
        """Create the file with its labels and an unset matrix; returns the SnpMemMap opened for writing (snpmemmap.py:88-123)."""
        assert order in ("F", "C"), "order must be 'F' or 'C'"
        dtype = np.dtype(dtype)
        row = np.array(iid, dtype=str).reshape(-1, 2)
        col = np.array(sid, dtype=str).reshape(-1)
        col_property = np.full((len(col), 3), np.nan) if pos is None else np.array(pos, dtype=np.float64).reshape(-1, 3)
        assert len(col_property) == len(col), "pos and sid must have the same length"
        with open(str(filename), "wb") as fp:
            np.save(fp, np.array([_MAGIC]))
            np.save(fp, np.array(["pstmemmap"]))
            np.save(fp, np.array([2]))
            np.save(fp, row)
            np.save(fp, col)
            np.save(fp, np.empty((len(row), 0)))
            np.save(fp, col_property)
            np.save(fp, np.array([dtype]))
            np.save(fp, np.array([order]))
            np.save(fp, np.array([None]))
            offset = fp.tell()
        self = SnpMemMap(filename)
        shape = (len(row), len(col))
        if shape[0] * shape[1] == 0:
            val = np.empty(shape, dtype=dtype, order=order)     # np.memmap cannot map zero bytes
        else:
            val = np.memmap(str(filename), offset=offset, dtype=dtype, mode="r+", order=order, shape=shape)
        self._attach(row, col, col_property, val, offset, dtype, order)
        return self

    def flush(self):
        """Flush the matrix to disk and close the mapping (it is reopened read-only on the next access)."""
        if self._ran_once:
            if isinstance(self._val, np.memmap):
                self._val.flush()
            self._val = None
            self._ran_once = False

    @staticmethod
    def write(filename, snpreader, standardizer=Identity(), order="A", dtype=None, block_size=None, num_threads=None):
        """Write ``snpreader`` (optionally standardized) in SnpMemMap format (snpmemmap.py:145-190).  A file-backed reader is
        streamed in SNP blocks through the GPU: decode + standardize fused, one PCIe crossing per block."""
        filename = str(filename)
        block_size = block_size or max(100_000 // max(1, snpreader.iid_count), 1)
        in_memory = hasattr(snpreader, "val")
        if in_memory:
            if _is_tensor(snpreader.val):                       # a device-resident SnpData: stage it on the host first
                snpreader = snpreader.read(order="A", dtype=np.dtype(str(snpreader.val.dtype).replace("torch.", "")))
            v = snpreader.val
            if order == "A":
                order = "F" if (v.flags["F_CONTIGUOUS"] and not v.flags["C_CONTIGUOUS"]) else "C"
            dtype = dtype or v.dtype
        else:
            order = "F" if order == "A" else order
            dtype = dtype or np.float64
        dtype = np.dtype(dtype)
        out = SnpMemMap.empty(iid=snpreader.iid, sid=snpreader.sid, filename=filename + ".temp", pos=snpreader.pos, order=order, dtype=dtype)
        if in_memory:
            standardizer.standardize(snpreader, num_threads=num_threads)
            out.val[:, :] = snpreader.val
        else:
            fused = isinstance(standardizer, Standardizer) and standardizer._device_spec() is not None and snpreader._can_fuse()
            for start in range(0, snpreader.sid_count, block_size):
                block = snpreader[:, start:start + block_size]
                if fused:
                    data = block.read(order=order, dtype=dtype, num_threads=num_threads, standardizer=standardizer)
                else:
                    data = block.read(order=order, dtype=dtype, num_threads=num_threads)
                    standardizer.standardize(data, num_threads=num_threads)
                out.val[:, start:start + data.sid_count] = data.val
        out.flush()
        if os.path.exists(filename):
            os.remove(filename)
        shutil.move(filename + ".temp", filename)
        return SnpMemMap(filename)

    # ---- reading --------------------------------------------------------------------------------------------------------
    def _read(self, iid_index_or_none, sid_index_or_none, order, dtype, force_python_only, view_ok, num_threads, to_device=False,
              _standardize=None, _out=None):
        """Whole matrix: the mapping itself when ``view_ok`` allows, else a copy.  A subset is gathered by the host from the mapped
        file (only the touched pages are read) -- this is file I/O, not the compute path."""
        if _standardize is not None or _out is not None:
            raise NotImplementedError("SnpMemMap.read takes no standardizer= / out=: use .read().standardize(...)")
        val = self.val
        dtype = np.dtype(dtype)
        if iid_index_or_none is None and sid_index_or_none is None:
            ok_order = order == "A" or val.flags["F_CONTIGUOUS" if order == "F" else "C_CONTIGUOUS"]
            if view_ok and ok_order and val.dtype == dtype:
                out = val
            else:
                out = np.array(val, dtype=dtype, order=order)
        else:
            sub = val
            if sid_index_or_none is not None:                  # columns first for F order (contiguous), rows first for C order
                if self._order == "F":
                    sub = sub[:, sid_index_or_none]
                    sub = sub if iid_index_or_none is None else sub[iid_index_or_none, :]
                else:
                    sub = sub if iid_index_or_none is None else sub[iid_index_or_none, :]
                    sub = sub[:, sid_index_or_none]
            else:
                sub = sub[iid_index_or_none, :]
            out = np.array(sub, dtype=dtype, order="F" if order in ("F", "A") else "C")
        if to_device:
            import torch
            return torch.from_numpy(np.ascontiguousarray(out)).cuda()
        return out

    def _read_kernel(self, standardizer, block_size=None, order="A", dtype=np.float64, force_python_only=False, view_ok=False,
                     return_trained=False, num_threads=None, to_device=False):
        """``K = sum over column blocks of X_b X_b^T`` (snpreader.py:651-655) with the mapped matrix streamed through the GPU:
        each block is standardized on the device (``pstb_standardize``; per-SNP statistics, so blocks are exact) and multiplied on
        the tensor cores (``pstb_float_kernel``, accumulating)."""
        _no_python_path(force_python_only)
        import torch
        from . import device
        dtype = np.dtype(dtype)
        val = self.val
        n, m = val.shape
        spec = standardizer._device_spec() if isinstance(standardizer, Standardizer) else None
        if spec is None and not isinstance(standardizer, Identity):
            raise NotImplementedError("read_kernel on the GPU supports Unit, Beta, their trained forms and Identity")
        stats_in = standardizer._trained_stats_for(self.sid) if spec is not None else None
        block = int(block_size) if block_size else max(1, (256 << 20) // max(1, n * 8))
        K = torch.zeros((n, n), dtype=torch.float32, device="cuda")
        stats = np.empty((m, 2), dtype=np.float64)
        for start in range(0, m, block):
            stop = min(m, start + block)
            v = torch.from_numpy(np.array(val[:, start:stop], dtype=np.float64 if val.dtype == np.float64 else np.float32, order="F").T).cuda().t()
            if spec is not None:
                st = device.standardize(v, spec, stats=None if stats_in is None else np.asarray(stats_in)[start:stop])
                stats[start:stop] = st.cpu().numpy()
            K = device.float_kernel(v, K=K, accumulate=start > 0, mirror=False)
        from . import _lib
        _lib.check(_lib.lib.pstb_mirror_lower(K.data_ptr(), n, n, torch.cuda.current_stream().cuda_stream))
        out = device.convert_kernel(K, dtype)
        result = out if to_device else _symmetric_to_host(out, order)
        if return_trained:
            trained = standardizer if spec is None else standardizer._make_trained(self.sid, stats.astype(dtype if dtype in (np.float32, np.float64) else np.float64))
            return result, trained
        return result
