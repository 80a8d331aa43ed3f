"""Generate the golden vectors under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container only (``python tests/golden/make_golden.py``): it imports the
reference's Python layer from /root/reference (read-only) with ``tests/golden/_refstub`` standing
in for the uninstalled ``bed_reader`` Rust wheel.  Every standardize / kernel golden is produced
by the reference's own code (``force_python_only=True`` -> ``standardizer.py:135-211``,
``snpreader.py:623-668``, ``snpdata.py:190-214``); every decode golden is checked against a file
the reference ships (``*.pst.npz``, ``toydata10.snp.npz``) before it is stored.  The small
``.bed/.bim/.fam`` inputs are copied beside the vectors so that the GPU box (which has no
/root/reference) can run the parity tests.
"""
import os
import shutil
import sys
import warnings
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
np.NAN = np.NaN = np.nan                      # reference predates NumPy 2 (util/__init__.py:329)
sys.path.insert(0, os.path.join(HERE, "_refstub"))
sys.path.insert(0, REF)
warnings.simplefilter("ignore")

import pysnptools.pstreader.pstreader as _pr   # noqa: E402


def _process_ndarray(indexer):                 # NumPy-2 fix for pstreader.py:636-645 (dtype=np.integer)
    if len(indexer) == 0:
        return np.zeros((0), dtype=np.int64)
    if indexer.dtype == bool:
        return np.arange(len(indexer), dtype=np.int64)[indexer]
    return indexer


_pr.PstReader._process_ndarray = staticmethod(_process_ndarray)

from pysnptools.snpreader import Bed, SnpData                    # noqa: E402
from pysnptools.standardizer import Unit, Beta, DiagKtoN         # noqa: E402
from pysnptools.kernelreader import SnpKernel                    # noqa: E402

FILES = {
    "n300": "tests/datasets/all_chr.maf0.001.N300",
    "toydata": "pysnptools/examples/toydata.5chrom",
    "dbx": "tests/datasets/distributed_bed_test1_X",
    "snpgen": "tests/datasets/snpgen",
    "gen1": "tests/datasets/generate/gen1",
    "gen4": "tests/datasets/generate/gen4",
}


def to_i8(val):
    out = np.where(np.isnan(val), -127, val).astype(np.int8)
    assert np.array_equal(np.where(out == -127, np.nan, out.astype(np.float64)), val, equal_nan=True)
    return out


def main():
    data = os.path.join(HERE, "data")
    os.makedirs(data, exist_ok=True)
    for key, stem in FILES.items():
        for ext in (".bed", ".bim", ".fam"):
            dst = os.path.join(data, key + ext)
            shutil.copyfile(os.path.join(REF, stem + ext), dst)
            os.chmod(dst, 0o644)

    shutil.copyfile(os.path.join(REF, "pysnptools/examples/toydata.phe"), os.path.join(data, "toydata.phe"))   # kernelreader/test.py:125
    os.chmod(os.path.join(data, "toydata.phe"), 0o644)

    # DistributedBed fixture (snpreader/distributedbed.py:296-311): pieces written with count_A1=True == distributed_bed_test1_X
    dst_dir = os.path.join(data, "distributed_bed_test1")
    if os.path.isdir(dst_dir):
        shutil.rmtree(dst_dir)
    shutil.copytree(os.path.join(REF, "tests/datasets/distributed_bed_test1"), dst_dir)
    for root_, _dirs, files in os.walk(dst_dir):
        os.chmod(root_, 0o755)
        for f_ in files:
            os.chmod(os.path.join(root_, f_), 0o644)

    g = {}
    # ---------------- decode goldens (values the reference ships) ----------------
    n300 = Bed(os.path.join(REF, FILES["n300"] + ".bed"), count_A1=False)
    v = n300.read().val
    for npz in ("tests/datasets/all_chr.maf0.001.N300.pst.npz", "tests/datasets/little.pst.npz"):
        ref_val = np.load(os.path.join(REF, npz))["val"]
        assert np.array_equal(ref_val, v), npz
    assert (v[0, 0], v[0, 2], v[0, 3]) == (2.0, 1.0, 2.0)      # snpreader.py doctests
    g["n300_decode_i8"] = to_i8(v)
    toy = Bed(os.path.join(REF, FILES["toydata"] + ".bed"), count_A1=False)
    tv = toy.read().val
    t10 = np.load(os.path.join(REF, "pysnptools/examples/toydata10.snp.npz"))["val"]
    assert np.array_equal(t10, tv[:, :10])
    g["toydata_decode_first10"] = t10
    g["toydata_decode_i8"] = to_i8(tv)
    # missing-value counts of the NaN-bearing fixtures as counted from the raw file bytes during the survey (SURVEY.md Appendix E;
    # consistent with the generator's missing_rate = .218, snpreader/snpgen.py:166-179): pins these decodes independently of the stub
    # decoder that produced them.  The reference's own TestSnpGen.test1 / TestDistributedBed.test1 (generator output == committed
    # files) run on the GPU shim in tests/test_gpu_reference_dropin.py.
    MISSING = {"dbx": 2182, "snpgen": 1144}
    for key in ("dbx", "snpgen", "gen1", "gen4"):
        bed = Bed(os.path.join(REF, FILES[key] + ".bed"), count_A1=False)
        if key in MISSING:
            assert int(np.isnan(bed.read().val).sum()) == MISSING[key], (key, int(np.isnan(bed.read().val).sum()))
        g[key + "_decode_i8"] = to_i8(bed.read().val)
        g[key + "_decode_A1_i8"] = to_i8(Bed(os.path.join(REF, FILES[key] + ".bed"), count_A1=True).read().val)
    # subset semantics through the reference's own indexer composition (pstreader/_subset.py)
    sub = n300[::-2, [5, 3, 3, -1]][1:40:3, :].read(order="C", dtype=np.float32).val
    g["n300_subset_rev_f32"] = sub
    g["n300_subset_rev_iid"] = np.arange(300)[::-2][1:40:3]
    g["n300_subset_rev_sid"] = np.array([5, 3, 3, 1014])

    # ---------------- standardize goldens: reference python twins, float64 ----------------
    def std(reader, s):
        d = reader.read(order="F", dtype=np.float64)
        d, trained = d.standardize(s, return_trained=True, force_python_only=True)
        return d.val, np.array(trained.stats, dtype=np.float64)

    for key, reader in (("n300", n300), ("dbx", Bed(os.path.join(REF, FILES["dbx"] + ".bed"), count_A1=False)),
                        ("snpgen", Bed(os.path.join(REF, FILES["snpgen"] + ".bed"), count_A1=False))):
        cols = slice(0, 160) if key == "n300" else slice(None)
        for name, s in (("unit", Unit()), ("beta_1_25", Beta(1, 25)), ("beta_2_10", Beta(2, 10))):
            val, stats = std(reader, s)
            g["{0}_{1}_val".format(key, name)] = val[:, cols]
            g["{0}_{1}_stats".format(key, name)] = stats
    assert "{0:.6f}".format(g["n300_unit_val"][0, 0]) == "0.229416"          # unit.py:18-20
    assert "{0:.6f}".format(g["n300_beta_1_25_val"][0, 0]) == "0.680802"     # beta.py:19-23

    # trained standardizer (standardizer.py:33-42, unittrained.py:19-30, betatrained.py:19-30)
    train_idx, test_idx = list(range(10, 300)), list(range(0, 10))
    tr, trained = Unit().standardize(n300[train_idx, :].read(), return_trained=True, force_python_only=True)
    te = n300[test_idx, :].read().standardize(trained, force_python_only=True)
    assert "{0:.6f}".format(tr.val[0, 0]) == "0.233550" and abs(te.val[0, 0] - 0.23354968324845735) < 1e-15
    g["n300_trained_unit_stats"] = np.array(trained.stats)
    g["n300_trained_unit_test_val"] = te.val
    trb, trainedb = Beta(1, 25).standardize(n300[train_idx, :].read(), return_trained=True, force_python_only=True)
    teb = n300[test_idx, :].read().standardize(trainedb, force_python_only=True)
    g["n300_trained_beta_stats"] = np.array(trainedb.stats)
    g["n300_trained_beta_test_val"] = teb.val

    # NaN / SNC injection as NaNCNCTestCases does (test.py:1296-1358)
    nc = n300[:, :64].read(order="C", dtype=np.float64)
    nc.val[0, 0] = np.nan
    nc.val[:, 1] = 2.0
    g["n300_nancnc_input"] = nc.val.copy()
    for name, s in (("unit", Unit()), ("beta_1_25", Beta(1, 25))):
        d = SnpData(iid=nc.iid, sid=nc.sid, val=nc.val.copy())
        d, trained = d.standardize(s, return_trained=True, force_python_only=True)
        assert d.val[0, 0] == 0 and np.all(d.val[:, 1] == 0)
        g["n300_nancnc_{0}_val".format(name)] = d.val
        g["n300_nancnc_{0}_stats".format(name)] = np.array(trained.stats)

    # ---------------- kernel goldens ----------------
    K = n300.read_kernel(Unit(), force_python_only=True).val
    assert "{0:.6f}".format(K[0, 0]) == "901.421836"                         # snpreader.py:308-313
    g["n300_unit_K"] = K
    Kb = n300.read_kernel(Unit(), block_size=100, force_python_only=True).val
    assert np.allclose(K, Kb, rtol=1e-12, atol=1e-9)
    g["n300_beta_1_25_K"] = n300.read_kernel(Beta(1, 25), block_size=500, force_python_only=True).val
    dbx = Bed(os.path.join(REF, FILES["dbx"] + ".bed"), count_A1=False)
    g["dbx_unit_K"] = dbx.read_kernel(Unit(), block_size=10, force_python_only=True).val
    Kt_ship = np.load(os.path.join(REF, "pysnptools/examples/toydata.kernel.npz"))["val"]
    Kt = toy.read_kernel(Unit(), force_python_only=True).val
    assert "{0:.6f}".format(Kt[0, 0]) == "9923.069928"                       # standardizer.py:26-28
    assert np.linalg.norm(Kt - Kt_ship) / np.linalg.norm(Kt_ship) < 1e-14
    g["toydata_unit_K_shipped"] = Kt_ship
    kd = SnpKernel(toy, Unit()).read().standardize(DiagKtoN())
    assert "{0:.6f}".format(kd.val[0, 0]) == "0.992307"                      # snpkernel.py:39-41
    g["toydata_unit_K_diagKtoN_00"] = np.array(kd.val[0, 0])
    # SnpKernel subset semantics (kernelreader/test.py:235-247): standardize on all iids, then slice
    g["n300_unit_K_every2"] = SnpKernel(n300, Unit())[::2].read(force_python_only=True).val

    np.savez_compressed(os.path.join(HERE, "golden.npz"), **g)
    print("wrote", os.path.join(HERE, "golden.npz"), os.path.getsize(os.path.join(HERE, "golden.npz")), "bytes;", len(g), "arrays")


if __name__ == "__main__":
    main()
