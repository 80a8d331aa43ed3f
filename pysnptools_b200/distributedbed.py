"""DistributedBed: a directory of per-chromosome-piece ``.bed`` files read as one SNP matrix.

Mirrors ``pysnptools/snpreader/distributedbed.py`` (reader 44-104, writer 107-207, pieces 209-268) for local directories:
``reader_name_list.npz`` names the pieces; every piece is a plain ``.bed/.bim/.fam`` written with ``count_A1=True``.
Pieces are SNP ranges -- exactly the per-GPU shard unit of the kinship path -- so ``read_kernel`` streams the pieces it
needs through the GPU one after the other and accumulates K, and a multi-GPU job can hand each rank its own pieces.
"""
import os

import numpy as np

from .snpreader import Bed, SnpReader, _kernel_chunk
from .standardizer import _no_python_path


class DistributedBed(SnpReader):
    def __init__(self, storage):
        super(DistributedBed, self).__init__()
        self._storage = str(storage)
        self._pieces = None

    def __repr__(self):
        return "{0}('{1}')".format(self.__class__.__name__, self._storage)

    def __getstate__(self):
        return self._storage

    def __setstate__(self, state):
        self.__init__(state)

    def _run_once(self):
        if self._pieces is not None:
            return
        names = np.array(np.load(os.path.join(self._storage, "reader_name_list.npz"))["reader_name_list"], dtype="str")
        self._pieces = [Bed(os.path.join(self._storage, str(n)), count_A1=True, skip_format_check=True) for n in names]
        self._row = self._pieces[0].row if self._pieces else np.empty((0, 2), dtype=str)
        counts = [p.sid_count for p in self._pieces]
        self._starts = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        self._col = np.concatenate([p.col for p in self._pieces]) if self._pieces else np.empty(0, dtype=str)
        self._col_property = np.concatenate([p.col_property for p in self._pieces]) if self._pieces else np.empty((0, 3))
        for p in self._pieces:
            assert np.array_equal(p.row, self._row), "all pieces must list the same individuals"

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

    def _split(self, sid_index_or_none):
        """(piece number, sid indices local to the piece, output columns) for every piece a selection touches, in piece order."""
        self._run_once()
        sid = np.arange(self.sid_count, dtype=np.int64) if sid_index_or_none is None else np.asarray(sid_index_or_none, dtype=np.int64)
        piece_of = np.searchsorted(self._starts, sid, side="right") - 1
        out = []
        for k in np.unique(piece_of):
            where = np.nonzero(piece_of == k)[0]
            out.append((int(k), sid[where] - self._starts[k], where))
        return sid, out

    def _read(self, iid_index_or_none, sid_index_or_none, order, dtype, force_python_only, view_ok, num_threads, to_device=False):
        _no_python_path(force_python_only)
        assert not to_device, "DistributedBed returns NumPy arrays; use its pieces for device results"
        dtype = np.dtype(dtype)
        sid, parts = self._split(sid_index_or_none)
        n_iid = self.iid_count if iid_index_or_none is None else len(iid_index_or_none)
        val = np.empty((n_iid, len(sid)), dtype=dtype, order="F" if order in ("F", "A") else "C")
        for k, local, where in parts:                 # only the pieces that hold requested SNPs are touched
            val[:, where] = self._pieces[k]._read(iid_index_or_none, local, order, dtype, False, view_ok, num_threads)
        return val

    def _root_and_indices(self):
        return self, None, None

    def _read_kernel_pieces(self, iid_idx, sid_idx, standardizer, block_size, dtype, return_trained):
        """K accumulated over the pieces that hold the selected SNPs (each piece = one packed store on the GPU)."""
        from . import device
        sid, parts = self._split(sid_idx)
        spec = standardizer._device_spec()
        stats_in = standardizer._trained_stats_for(self.sid[sid])
        if spec is None:
            spec, stats_in = ("unit",), np.tile(np.array([[0.0, 1.0]]), (len(sid), 1))
        K = None
        stats = np.empty((len(sid), 2), dtype=np.float64)
        n_iid = self.iid_count if iid_idx is None else len(iid_idx)
        low_term = device.low_term_for(len(sid), n_iid, spec)     # the precision mode follows the SNP count of the whole kernel
        for k, local, where in parts:
            piece = self._pieces[k]
            store, _ = piece._store_for(None)
            K, st = device.snp_kernel(store, iid_idx, local, count_A1=True, standardizer=spec,
                                      stats=None if stats_in is None else np.asarray(stats_in)[where],
                                      chunk=_kernel_chunk(block_size, n_iid, len(local)), K=K, accumulate=K is not None, mirror=False,
                                      low_term=low_term)
            stats[where] = st.cpu().numpy()
            piece._device_store = None                # one piece resident at a time
        import torch
        from . import _lib
        if K is None:
            K = torch.zeros((n_iid, n_iid), dtype=torch.float32, device="cuda")
        _lib.check(_lib.lib.pstb_mirror_lower(K.data_ptr(), n_iid, n_iid, torch.cuda.current_stream().cuda_stream))
        out = device.convert_kernel(K, dtype).cpu().numpy()
        return (out, standardizer._make_trained(self.sid[sid], stats.astype(dtype))) if return_trained else out

    @staticmethod
    def write(storage, snpreader, piece_per_chrom_count=1, updater=None, runner=None):
        """Write ``snpreader`` as chromosome pieces (``chrom{c}.piece{p}of{n}.bed``, count_A1=True) + ``reader_name_list.npz``."""
        storage = str(storage)
        os.makedirs(storage, exist_ok=True)
        chrom = snpreader.pos[:, 0]
        for c in sorted(set(chrom.tolist()), key=lambda x: (x != x, x)):
            # distributedbed.py:167-169: every chromosome must be an integer (a NaN chromosome would be dropped silently otherwise)
            assert c == c and c == int(c), "DistributedBed.write expects all chromosomes to be integers (not '{0}')".format(c)
        names = []
        for c in sorted(set(chrom)):
            idx = np.nonzero(chrom == c)[0]
            for p in range(piece_per_chrom_count):
                lo, hi = len(idx) * p // piece_per_chrom_count, len(idx) * (p + 1) // piece_per_chrom_count
                if hi <= lo and piece_per_chrom_count > len(idx):
                    continue
                name = "chrom{0}.piece{1}of{2}.bed".format(int(c), p, piece_per_chrom_count)
                trio = [os.path.join(storage, name[:-3] + ext) for ext in ("bim", "fam", "bed")]
                present = [os.path.exists(f) for f in trio]
                if sum(present) < 3:                # a piece is skipped only when all three files exist; leftovers of a partial write are removed
                    for f, there in zip(trio, present):
                        if there:
                            os.remove(f)
                    Bed.write(trio[-1], snpreader[:, idx[lo:hi]].read(dtype=np.float64), count_A1=True)
                names.append(name)
        np.savez(os.path.join(storage, "reader_name_list.npz"), reader_name_list=np.array(names, dtype="S"))
        # metadata.npz: the cache file the reference's DistributedBed opens unconditionally (its _MergeSIDs cache, snpreader/_mergesids.py:9-23:
        # row / row_property / col / col_property of the merged reader + the SNP count of every piece), so a directory written here reads there
        out = DistributedBed(storage)
        out._run_once()
        np.savez(os.path.join(storage, "metadata.npz"), _row=np.array(out._row, dtype="S"), _row_property=np.empty((len(out._row), 0)),
                 _col=np.array(out._col, dtype="S"), _col_property=out._col_property,
                 sid_count_list=np.array([p.sid_count for p in out._pieces], dtype=np.int32))     # snpreader/_mergesids.py:9-23
        return out
