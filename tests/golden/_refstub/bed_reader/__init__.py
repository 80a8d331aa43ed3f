"""Stand-in for the third-party ``bed_reader`` wheel, backed by ``oracle/bed_oracle.py``.

TEST INFRASTRUCTURE: used only by ``tests/golden/make_golden.py`` so that the reference's own
Python layer (``/root/reference/pysnptools``) can be imported in the build container, where
the Rust wheel is not installed.  The nine symbols are the ones the reference imports
(SURVEY.md section 0, item 2).
"""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "..", ".."))
from oracle import bed_oracle as _o  # noqa: E402


def get_num_threads(num_threads=None):
    if num_threads is not None:
        return num_threads
    for k in ("PST_NUM_THREADS", "NUM_THREADS", "MKL_NUM_THREADS"):
        if k in os.environ:
            return int(os.environ[k])
    return os.cpu_count()


class open_bed(object):
    _FAM = {"fid": 0, "iid": 1, "father": 2, "mother": 3, "sex": 4, "pheno": 5}
    _BIM = {"chromosome": 0, "sid": 1, "cm_position": 2, "bp_position": 3, "allele_1": 4, "allele_2": 5}

    def __init__(self, filepath, iid_count=None, sid_count=None, properties={}, count_A1=True,
                 num_threads=None, skip_format_check=False, fam_filepath=None, bim_filepath=None):
        self.filepath = str(filepath)
        self.count_A1 = count_A1
        self.properties = dict(properties)
        self.skip_format_check = skip_format_check
        base = self.filepath[:-4] if self.filepath.endswith(".bed") else self.filepath
        self.fam_filepath = fam_filepath or base + ".fam"
        self.bim_filepath = bim_filepath or base + ".bim"
        self._cache = {}

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def _column(self, name):
        if name in self._cache:
            return self._cache[name]
        if self.properties.get(name, 0) is not None and name in self.properties:
            val = np.asarray(self.properties[name])
        else:
            table, path = (self._FAM, self.fam_filepath) if name in self._FAM else (self._BIM, self.bim_filepath)
            col = table[name]
            with open(path) as f:
                items = [line.split()[col] for line in f if line.strip()]
            if name == "cm_position":
                val = np.array(items, dtype=np.float64)
            elif name == "bp_position":
                val = np.array(items, dtype=np.float64).astype(np.int64)
            else:
                val = np.array(items, dtype=str)
        self._cache[name] = val
        return val

    fid = property(lambda s: s._column("fid"))
    iid = property(lambda s: s._column("iid"))
    sid = property(lambda s: s._column("sid"))
    chromosome = property(lambda s: s._column("chromosome"))
    cm_position = property(lambda s: s._column("cm_position"))
    bp_position = property(lambda s: s._column("bp_position"))

    @property
    def iid_count(self):
        return len(self.iid) if ("iid" in self.properties and self.properties["iid"] is not None) else _o.count_lines(self.fam_filepath)

    @property
    def sid_count(self):
        return len(self.sid) if ("sid" in self.properties and self.properties["sid"] is not None) else _o.count_lines(self.bim_filepath)

    def read(self, index=None, dtype="float32", order="F", force_python_only=False, num_threads=None):
        n, m = self.iid_count, self.sid_count
        packed = _o.read_packed(self.filepath, n, m, self.skip_format_check)
        ii, si = (None, None) if index is None else index
        return _o.decode(packed, n, ii, si, self.count_A1, np.dtype(dtype), order)


def to_bed(*a, **k):
    raise NotImplementedError("writing is outside the golden-vector generator")


def _std(snps, is_beta, a, b, apply_in_place, use_stats, stats, num_threads):
    out, st = _o.standardize(snps, is_beta, a, b, use_stats, stats)
    if not use_stats:
        stats[...] = st
    if apply_in_place:
        snps[...] = out


standardize_f64 = _std
standardize_f32 = _std


def _subset(val, row, col, out, num_threads):
    out[...] = val[np.asarray(row, dtype=np.int64)][:, np.asarray(col, dtype=np.int64)]


subset_f64_f64 = subset_f32_f64 = subset_f32_f32 = _subset
