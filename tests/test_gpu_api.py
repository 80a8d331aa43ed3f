"""GPU: the reference-facing API (Bed / SnpData / Unit / Beta / SnpKernel / bed_reader shim) against goldens + oracle.

Reads like the reference's own tests (pysnptools/test.py, kernelreader/test.py) for the hot path.
"""
import os
import pickle
import sys
import warnings

import numpy as np
import pytest

from conftest import DATA_DIR, ROOT, SHAPES, fixture_packed, i8_to_float

pytestmark = pytest.mark.gpu
warnings.simplefilter("ignore", DeprecationWarning)


def _bed(name, **kw):
    from pysnptools_b200 import Bed
    kw.setdefault("count_A1", False)
    return Bed(os.path.join(DATA_DIR, name + ".bed"), **kw)


def rel_fro(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("name", list(SHAPES))
def test_bed_read_respects_inputs(name, golden):
    """test_respect_read_inputs (test.py:912-1005) + test_bed_int8 (288-324): dtype / order contract and values."""
    want = golden[name + "_decode_i8"]
    bed = _bed(name)
    for order in ("F", "C", "A"):
        for dtype in (np.float32, np.float64):
            d = bed.read(order=order, dtype=dtype)
            assert d.val.dtype == dtype and d.val.flags["C_CONTIGUOUS" if order == "C" else "F_CONTIGUOUS"]
            assert np.array_equal(d.val, i8_to_float(want, dtype), equal_nan=True)
            assert d.iid.shape == (want.shape[0], 2) and len(d.sid) == want.shape[1]
        d8 = bed.read(order=order, dtype="int8", _require_float32_64=False)
        assert d8.val.dtype == np.int8 and np.array_equal(d8.val, want)
    a1 = _bed(name, count_A1=True).read(dtype="int8", _require_float32_64=False).val
    assert np.array_equal(a1, golden[name + "_decode_A1_i8"]) if name + "_decode_A1_i8" in golden.files else True
    obs = want != -127
    assert np.array_equal(a1[obs], 2 - want[obs])                                   # test.py:226-232: A1 == 2 - A2


def test_bed_subsets(golden, oracle):
    bed = _bed("n300")
    packed, n, m = fixture_packed("n300")
    sub = bed[::-2, [5, 3, 3, -1]][1:40:3, :].read(order="C", dtype=np.float32)
    assert np.array_equal(sub.val, golden["n300_subset_rev_f32"], equal_nan=True)
    assert np.array_equal(sub.sid, bed.sid[[5, 3, 3, 1014]])
    rev = bed[::-1, ::-3].read()                                                    # reversed strides (test.py:812-862)
    assert np.array_equal(rev.val, oracle.decode(packed, n, np.arange(n)[::-1], np.arange(m)[::-3]), equal_nan=True)
    mask = np.arange(n) % 7 == 0
    assert np.array_equal(bed[mask, 2:5].read().val, oracle.decode(packed, n, np.nonzero(mask)[0], [2, 3, 4]), equal_nan=True)
    one = bed[np.int64(3), 4].read()                                                # test_scalar_index (test.py:326-331)
    assert one.val.shape == (1, 1)
    assert bed[[], :].read().val.shape == (0, m) and bed[:, []].read().val.shape == (n, 0)
    with pytest.raises(IndexError):
        bed[[n], :]
    dev = bed[:, :10].read(dtype=np.float32, to_device=True)
    assert dev.val.is_cuda and np.array_equal(dev.val.cpu().numpy(), oracle.decode(packed, n, None, np.arange(10), dtype=np.float32))
    clone = pickle.loads(pickle.dumps(bed))
    assert np.array_equal(clone[:5, :5].read().val, bed[:5, :5].read().val)


def test_standardize_unit_and_beta(golden):
    """test_standardize_bed (test.py:582-640), doctests unit.py:18-20 / beta.py:19-23 / unittrained.py:19-30."""
    from pysnptools_b200 import Beta, Unit
    bed = _bed("n300")
    for tag, s in (("unit", Unit()), ("beta_1_25", Beta(1, 25)), ("beta_2_10", Beta(2, 10))):
        want, wst = golden["n300_{0}_val".format(tag)], golden["n300_{0}_stats".format(tag)]
        for order in ("F", "C"):
            for dtype, tol in ((np.float64, 1e-11), (np.float32, 1e-5)):
                d = bed.read(order=order, dtype=dtype)
                out, trained = d.standardize(s, return_trained=True)
                assert out is d and not np.isnan(d.val).any()
                np.testing.assert_allclose(d.val[:, :want.shape[1]], want, rtol=tol, atol=tol)
                np.testing.assert_allclose(trained.stats, wst, rtol=1e-12 if dtype == np.float64 else 1e-6)
                assert trained.stats.dtype == dtype and trained.is_constant
                arr = bed.read(order=order, dtype=dtype).val                         # bare ndarray (deprecated but used by the tests)
                s.standardize(arr)
                assert np.array_equal(arr, d.val)
    d = bed.read().standardize(Unit())
    assert "{0:.6f}".format(d.val[0, 0]) == "0.229416" and repr(d).endswith("Unit())")
    assert "{0:.6f}".format(bed.read().standardize(Beta(1, 25)).val[0, 0]) == "0.680802"
    train, trained = bed[10:, :].read().standardize(Unit(), return_trained=True)
    assert "{0:.6f}".format(train.val[0, 0]) == "0.233550"
    test = bed[:10, :].read().standardize(trained)
    np.testing.assert_allclose(test.val, golden["n300_trained_unit_test_val"], rtol=1e-11, atol=1e-13)
    testb = bed[:10, :].read().standardize(bed[10:, :].read().standardize(Beta(1, 25), return_trained=True)[1])
    np.testing.assert_allclose(testb.val, golden["n300_trained_beta_test_val"], rtol=1e-11, atol=1e-13)
    # trained standardizer re-indexed by sid
    part = bed[:10, [7, 2]].read().standardize(trained)
    np.testing.assert_allclose(part.val, golden["n300_trained_unit_test_val"][:, [7, 2]], rtol=1e-11, atol=1e-13)


def test_nan_and_snc_cases(golden):
    """NaNCNCTestCases (test.py:1296-1358) and kernelreader test_cpp_std (56-110)."""
    from pysnptools_b200 import Beta, SnpData, Unit
    x = golden["n300_nancnc_input"]
    iid = [["0", "i{0}".format(k)] for k in range(x.shape[0])]
    sid = ["s{0}".format(k) for k in range(x.shape[1])]
    for tag, s in (("unit", Unit()), ("beta_1_25", Beta(1, 25))):
        for dtype, rtol in ((np.float64, 1e-12), (np.float32, 1e-4)):
            for order in ("C", "F"):
                d = SnpData(iid=iid, sid=sid, val=np.array(x, dtype=dtype, order=order))
                d, trained = d.standardize(s, return_trained=True)
                assert d.val[0, 0] == 0 and np.all(d.val[:, 1] == 0) and np.isinf(trained.stats[1, 1])
                np.testing.assert_allclose(d.val, golden["n300_nancnc_{0}_val".format(tag)], rtol=rtol, atol=1e-6 if dtype == np.float32 else 1e-13)
                again = SnpData(iid=iid, sid=sid, val=np.array(x, dtype=dtype, order=order)).standardize(trained)
                np.testing.assert_allclose(again.val, d.val, rtol=1e-12 if dtype == np.float64 else 1e-5, atol=1e-14 if dtype == np.float64 else 1e-6)   # f32 stats are rounded, as in the reference


def test_snp_kernel(golden):
    """doctests snpreader.py:308-313 / snpkernel.py:39-41, test_merge_std, test_respect_inputs, test_subset, test_some_std."""
    from pysnptools_b200 import Beta, DiagKtoN, Identity, SnpKernel, Unit
    bed = _bed("n300")
    kd = bed.read_kernel(Unit())
    assert kd.val.dtype == np.float64 and kd.iid_count == 300 and np.array_equal(kd.iid, bed.iid)
    assert rel_fro(kd.val, golden["n300_unit_K"]) < 1e-5 and abs(kd.val[0, 0] - 901.421836) < 1e-2
    for block_size in (None, 1, 100, 500):
        for order in ("F", "C", "A"):
            for dtype in (np.float64, np.float32):
                k = SnpKernel(bed, Unit(), block_size=block_size).read(order=order, dtype=dtype)
                assert k.val.dtype == dtype and k.val.flags["F_CONTIGUOUS" if order == "F" else "C_CONTIGUOUS"]
                assert rel_fro(k.val.astype(np.float64), golden["n300_unit_K"]) < 1e-5
    assert rel_fro(bed.read_kernel(Beta(1, 25), block_size=500).val, golden["n300_beta_1_25_K"]) < 1e-5
    assert rel_fro(_bed("dbx").read_kernel(Unit(), block_size=10).val, golden["dbx_unit_K"]) < 1e-5
    toy = SnpKernel(_bed("toydata"), Unit()).read()
    assert rel_fro(toy.val, golden["toydata_unit_K_shipped"]) < 1e-5
    toy.standardize(DiagKtoN())
    assert abs(np.trace(toy.val) - 500) < 1e-6 and abs(toy.val[0, 0] - float(golden["toydata_unit_K_diagKtoN_00"])) < 1e-5
    every2 = SnpKernel(bed, Unit())[::2].read()                                      # standardize on all iids, then slice
    assert rel_fro(every2.val, golden["n300_unit_K_every2"]) < 1e-5 and every2.iid_count == 150
    # in-memory SnpData kernel (val.dot(val.T)) == kernel from the file (test_some_std, test.py:530-553)
    sd = bed.read().standardize(Unit())
    k_mem = sd.read_kernel(Identity())
    assert rel_fro(k_mem.val, golden["n300_unit_K"]) < 1e-5
    k_mem2 = bed.read(dtype=np.float32, order="C").read_kernel(Unit())
    assert rel_fro(k_mem2.val, golden["n300_unit_K"]) < 1e-5
    kernel, snp_trained, kernel_trained = SnpKernel(bed, Unit(), block_size=300)._read_with_standardizing(True, return_trained=True)
    np.testing.assert_allclose(snp_trained.stats, golden["n300_unit_stats"], rtol=1e-12)
    assert abs(np.trace(kernel.val) - 300) < 1e-6 and abs(kernel_trained.factor - 300 / np.trace(golden["n300_unit_K"])) < 1e-7
    # trained (constant) standardizer: subset pushed into the reader
    ktr = SnpKernel(bed, snp_trained)[:10].read()
    want = golden["n300_trained_unit_test_val"]
    _, full_trained = bed[10:, :].read().standardize(Unit(), return_trained=True)
    ktr2 = SnpKernel(bed, full_trained)[:10].read()
    assert rel_fro(ktr2.val, want @ want.T) < 1e-5 and ktr.val.shape == (10, 10)


def test_train_test_kernel(golden, oracle):
    """SnpKernel(train, standardizer, test=test): X_train X_test^T with the statistics learned on the train iids -- the product
    FaST-LMM predicts with; values pinned by the reference's trained-standardizer goldens (unittrained.py / betatrained.py)."""
    from pysnptools_b200 import Beta, SnpKernel, Unit
    bed = _bed("n300")
    packed, n, m = fixture_packed("n300")
    for std, key, args in ((Unit(), "unit", {}), (Beta(1, 25), "beta", dict(is_beta=True, a=1, b=25))):
        train, test = bed[10:, :], bed[:10, :]
        xr, st = oracle.standardize(oracle.decode(packed, n, np.arange(10, n)), **args)
        want = xr @ golden["n300_trained_{0}_test_val".format(key)].T                # test values standardized BY THE REFERENCE
        k = SnpKernel(train, std, test=test)
        assert k.iid0_count == n - 10 and k.iid1_count == 10 and k.shape == (n - 10, 10)
        kd = k.read()
        assert kd.val.shape == (n - 10, 10) and kd.val.dtype == np.float64 and kd.val.flags["F_CONTIGUOUS"]
        assert np.array_equal(kd.iid0, train.iid) and np.array_equal(kd.iid1, test.iid)
        assert rel_fro(kd.val, want) < 1e-5, rel_fro(kd.val, want)
        # a trained standardizer gives the same matrix; float32 / C order / block_size / column subset respected
        _, trained = train.read().standardize(std, return_trained=True)
        k32 = SnpKernel(train, trained, block_size=200, test=test).read(order="C", dtype=np.float32)
        assert k32.val.dtype == np.float32 and k32.val.flags["C_CONTIGUOUS"] and rel_fro(k32.val.astype(np.float64), want) < 1e-5
        sub = k[::3, [7, 2]].read()
        assert sub.val.shape == (len(range(0, n - 10, 3)), 2) and rel_fro(sub.val, want[::3][:, [7, 2]]) < 1e-5
        # SNPs are paired by name: a test reader with its SNPs in another order gives the same kernel
        perm = np.random.default_rng(0).permutation(m)
        assert rel_fro(SnpKernel(train, std, test=test[:, perm]).read().val, want) < 1e-5
    assert "test=" in repr(SnpKernel(bed[10:, :], Unit(), test=bed[:10, :]))


def test_bed_write_round_trips(tmp_path, oracle):
    """test_write_bed_f64cpp_* / test_write_x_x_cpp (test.py:671-765): 0/1/2/5 iids, NaN, both count_A1, illegal values."""
    from pysnptools_b200 import Bed, SnpData
    for n in (0, 1, 2, 5, 190):
        for m in (0, 3, 20):
            packed = oracle.synth_packed(n, 0, m, 0.2, seed=n + m) if n and m else np.zeros((m, (n + 3) // 4), np.uint8)
            val = oracle.decode(packed, n) if n and m else np.zeros((n, m))
            sd = SnpData(iid=[["f", "i{0}".format(k)] for k in range(n)], sid=["s{0}".format(k) for k in range(m)], val=val,
                         pos=[[1, 0.5, 100 + k] for k in range(m)])
            for a1 in (False, True):
                back = Bed.write(str(tmp_path / "w_{0}_{1}_{2}".format(n, m, a1)), sd, count_A1=a1)
                got = back.read()
                assert got.val.shape == (n, m) and np.array_equal(got.val, val, equal_nan=True)
                assert np.array_equal(got.iid, sd.iid) and np.array_equal(got.sid, sd.sid) and np.array_equal(got.pos, sd.pos)
    bad = SnpData(iid=[["f", "a"], ["f", "b"]], sid=["s"], val=np.array([[5.0], [1.0]]))
    with pytest.raises(ValueError):
        Bed.write(str(tmp_path / "bad"), bad, count_A1=False)


def test_bed_reader_shim(golden, oracle, tmp_path):
    """The nine bed_reader symbols as the reference calls them (bed.py:337-343, standardizer.py:114,120, util/__init__.py:341-375)."""
    sys.path.insert(0, os.path.join(ROOT, "pysnptools_b200", "compat"))
    try:
        import bed_reader
    finally:
        sys.path.pop(0)
    packed, n, m = fixture_packed("dbx")
    ob = bed_reader.open_bed(os.path.join(DATA_DIR, "dbx.bed"), properties={}, count_A1=False, num_threads=None, skip_format_check=False)
    ii, si = np.arange(n, dtype=np.uintp)[::-1], np.array([3, 1, 99], dtype=np.uintp)
    for order in ("F", "C"):
        for dtype in (np.float32, np.float64, np.int8):
            val = ob.read(index=(ii, si), order=order, dtype=dtype, force_python_only=False, num_threads=4)
            assert val.flags["F_CONTIGUOUS" if order == "F" else "C_CONTIGUOUS"]
            assert np.array_equal(val, oracle.decode(packed, n, ii, si, False, dtype), equal_nan=dtype != np.int8)
    assert np.array_equal(ob.read(index=(None, None), dtype=np.int8), golden["dbx_decode_i8"])
    raw = oracle.decode(packed, n)
    for fn, dtype in ((bed_reader.standardize_f64, np.float64), (bed_reader.standardize_f32, np.float32)):
        for order in ("F", "C"):
            snps = np.array(raw, dtype=dtype, order=order)
            stats = np.empty((m, 2), dtype=dtype, order=order)
            fn(snps, True, 1.0, 25.0, True, False, stats, 2)
            np.testing.assert_allclose(snps, golden["dbx_beta_1_25_val"], rtol=1e-11 if dtype == np.float64 else 1e-5, atol=1e-6 if dtype == np.float32 else 1e-13)
            np.testing.assert_allclose(stats, golden["dbx_beta_1_25_stats"], rtol=1e-12 if dtype == np.float64 else 1e-6)
            again = np.array(raw, dtype=dtype, order=order)
            fn(again, True, 1.0, 25.0, True, True, stats, 2)
            np.testing.assert_allclose(again, snps, rtol=1e-6, atol=1e-7)
    v3 = np.random.default_rng(0).normal(size=(9, 7, 3))
    rows, cols = np.array([8, 0, 3], dtype=np.uintp), np.array([6, 6, 1, 0], dtype=np.uintp)
    out = np.full((3, 4, 3), np.nan, order="F")
    bed_reader.subset_f64_f64(np.asfortranarray(v3), rows, cols, out, 1)
    assert np.array_equal(out, v3[rows][:, cols])
    out32 = np.full((3, 4, 3), np.nan, dtype=np.float64)
    bed_reader.subset_f32_f64(v3.astype(np.float32), rows, cols, out32, 1)
    assert np.array_equal(out32, v3.astype(np.float32)[rows][:, cols].astype(np.float64))
    path = str(tmp_path / "shim.bed")
    bed_reader.to_bed(path, raw, properties={"fid": ["f"] * n, "iid": ["i%d" % k for k in range(n)], "sid": ["s%d" % k for k in range(m)]}, count_A1=False)
    back = bed_reader.open_bed(path, count_A1=False)
    assert np.array_equal(back.read(dtype=np.float64), raw, equal_nan=True) and back.sid[3] == "s3"
    with pytest.raises(ValueError):
        bed_reader.to_bed(str(tmp_path / "bad.bed"), np.full((3, 2), 7.0), count_A1=False)


def test_util_sub_matrix(oracle):
    """test_sub_matrix (util/test.py:118-128) and PstReader test_every_read (pstreader/test.py:118-133)."""
    from pysnptools_b200.util import sub_matrix
    rng = np.random.default_rng(1)
    for shape in ((11, 9), (11, 9, 1), (11, 9, 3)):
        src = rng.normal(size=shape)
        rows, cols = np.arange(10, -1, -2), np.arange(8, -1, -1)
        for order_from in ("C", "F"):
            for order_to in ("C", "F"):
                for dt_from, dt_to in ((np.float64, np.float64), (np.float32, np.float64), (np.float32, np.float32)):
                    a = np.array(src, dtype=dt_from, order=order_from)
                    out = sub_matrix(a, rows, cols, order=order_to, dtype=dt_to)
                    assert out.dtype == dt_to and out.flags["C_CONTIGUOUS" if order_to == "C" else "F_CONTIGUOUS"]
                    assert np.array_equal(out, a[rows][:, cols].astype(dt_to))


def test_fused_read_standardize_extension(golden):
    """read(standardizer=...) == read().standardize(...) in one GPU pass (host and device results, trained reuse)."""
    from pysnptools_b200 import Beta, Unit
    bed = _bed("n300")
    for tag, s in (("unit", Unit()), ("beta_1_25", Beta(1, 25))):
        want, wst = golden["n300_{0}_val".format(tag)], golden["n300_{0}_stats".format(tag)]
        for order in ("F", "C"):
            d, trained = bed.read(order=order, dtype=np.float64, standardizer=s, return_trained=True)
            assert d.val.flags["F_CONTIGUOUS" if order == "F" else "C_CONTIGUOUS"]
            np.testing.assert_allclose(d.val[:, :want.shape[1]], want, rtol=1e-11, atol=1e-13)
            np.testing.assert_allclose(trained.stats, wst, rtol=1e-12)
        sub = bed[::-2, 5:40]
        two_step = sub.read(dtype=np.float32).standardize(s).val
        fused = sub.read(dtype=np.float32, standardizer=s).val
        assert np.array_equal(fused, two_step) or np.allclose(fused, two_step, rtol=1e-6, atol=1e-7)
        dev_val = sub.read(dtype=np.float32, standardizer=s, to_device=True).val
        assert dev_val.is_cuda and np.array_equal(dev_val.cpu().numpy(), fused)
    _, trained = bed[10:, :].read(standardizer=Unit(), return_trained=True)
    test = bed[:10, :].read(standardizer=trained)
    np.testing.assert_allclose(test.val, golden["n300_trained_unit_test_val"], rtol=1e-11, atol=1e-13)


def test_distributed_bed(golden, tmp_path):
    """TestDistributedBed.test1 (snpreader/distributedbed.py:296-311): pieces == distributed_bed_test1_X; K over pieces."""
    from pysnptools_b200 import DistributedBed, Unit
    d = DistributedBed(os.path.join(DATA_DIR, "distributed_bed_test1"))
    want = i8_to_float(golden["dbx_decode_i8"])
    for order in ("F", "C"):
        assert np.array_equal(d.read(order=order).val, want, equal_nan=True)
    sub = d[::-3, [99, 0, 50, 1, 98]].read(dtype=np.float32)
    assert np.array_equal(sub.val, want[::-3][:, [99, 0, 50, 1, 98]].astype(np.float32), equal_nan=True)
    K, trained = d._read_kernel(Unit(), block_size=10, return_trained=True)
    assert rel_fro(K, golden["dbx_unit_K"]) < 1e-5
    np.testing.assert_allclose(trained.stats, golden["dbx_unit_stats"], rtol=1e-12)
    assert rel_fro(d.read_kernel(Unit()).val, golden["dbx_unit_K"]) < 1e-5
    back = DistributedBed.write(str(tmp_path / "dist"), _bed("dbx"), piece_per_chrom_count=2)
    assert back.sid_count == 100 and np.array_equal(back.read().val, want, equal_nan=True)
    assert np.array_equal(back.sid, d.sid)
    buf = np.empty((100, 100), order="F")
    assert d.read(out=buf).val is buf and np.array_equal(buf, want, equal_nan=True)          # out= on a reader without a fused path
    assert os.path.exists(str(tmp_path / "dist" / "metadata.npz"))
    os.remove(str(tmp_path / "dist" / "chrom1.piece0of2.fam"))                                # a partial piece is rewritten, a complete one skipped
    DistributedBed.write(str(tmp_path / "dist"), _bed("dbx"), piece_per_chrom_count=2)
    assert os.path.exists(str(tmp_path / "dist" / "chrom1.piece0of2.fam"))
    bad = _bed("dbx").read()
    bad.pos[3, 0] = np.nan
    with pytest.raises(AssertionError, match="integers"):
        DistributedBed.write(str(tmp_path / "bad"), bad)


def test_distributed_bed_pieces_as_rank_shards(golden):
    """Multi-GPU DistributedBed path with the ranks run one after the other on this GPU: every piece goes to exactly one rank,
    the partial kernels sum to the golden K, the gathered statistics come back in SNP order; world 1 == read_kernel_multi_gpu."""
    import torch
    from pysnptools_b200 import DistributedBed
    from pysnptools_b200.parallel import assign_pieces, distributed_bed_partial_kernel, read_kernel_multi_gpu
    d = DistributedBed(os.path.join(DATA_DIR, "distributed_bed_test1"))
    for world in (1, 2, 3):
        K = torch.zeros((100, 100), dtype=torch.float32, device="cuda")
        stats = np.full((100, 2), np.nan)
        seen = []
        for rank in range(world):
            K_r, st_r, where = distributed_bed_partial_kernel(d, rank, world, chunk=64)
            K += K_r
            stats[where] = st_r.cpu().numpy()
            seen += list(where)
        assert sorted(seen) == list(range(100))
        Kl = np.tril(K.double().cpu().numpy())
        full = Kl + np.tril(Kl, -1).T
        assert rel_fro(full, golden["dbx_unit_K"]) < 1e-5
        np.testing.assert_allclose(stats, golden["dbx_unit_stats"], rtol=1e-12)
    K1, st1 = read_kernel_multi_gpu(d)                                                  # no process group: world 1
    assert rel_fro(K1.double().cpu().numpy(), golden["dbx_unit_K"]) < 1e-5
    np.testing.assert_allclose(st1.cpu().numpy(), golden["dbx_unit_stats"], rtol=1e-12)
    d._run_once()
    assert sum(len(o) for o in assign_pieces([p.sid_count for p in d._pieces], 3)) == len(d._pieces)


def test_snpmemmap_staging(golden, tmp_path):
    """SnpMemMap as host-side staging: Bed -> (fused decode + standardize per SNP block) -> mapped file; kernels streamed from a mapped file."""
    from pysnptools_b200 import Beta, Identity, SnpKernel, SnpMemMap, Unit
    bed = _bed("n300")
    raw = SnpMemMap.write(str(tmp_path / "raw.snp.memmap"), bed, block_size=333)                    # Identity: the decoded matrix
    assert isinstance(raw.val, np.memmap) and raw.val.dtype == np.float64 and raw.val.flags["F_CONTIGUOUS"]
    assert np.array_equal(raw.val, i8_to_float(golden["n300_decode_i8"]), equal_nan=True)
    assert np.array_equal(raw.sid, bed.sid) and np.array_equal(raw.iid, bed.iid) and np.array_equal(raw.pos, bed.pos, equal_nan=True)
    for std, key in ((Unit(), "unit"), (Beta(1, 25), "beta_1_25")):
        for order, dtype, tol in (("F", np.float64, 1e-9), ("C", np.float32, 2e-6)):
            sm = SnpMemMap.write(str(tmp_path / "std.snp.memmap"), bed, standardizer=std, order=order, dtype=dtype, block_size=100)
            want = golden["n300_{0}_val".format(key)]
            assert sm.val.dtype == dtype and sm.val.flags["F_CONTIGUOUS" if order == "F" else "C_CONTIGUOUS"]
            assert np.max(np.abs(sm.val[:, : want.shape[1]] - want)) <= tol * max(1.0, np.max(np.abs(want)))
    # kernels streamed from mapped files: raw values standardized on the device block by block, or already standardized values
    K, trained = raw._read_kernel(Unit(), block_size=300, return_trained=True)
    assert rel_fro(K, golden["n300_unit_K"]) < 1e-5 and np.array_equal(K, K.T)
    np.testing.assert_allclose(trained.stats, golden["n300_unit_stats"], rtol=1e-12)
    assert rel_fro(SnpKernel(raw, Beta(1, 25), block_size=500).read().val, golden["n300_beta_1_25_K"]) < 1e-5
    sm = SnpMemMap.write(str(tmp_path / "unit.snp.memmap"), bed, standardizer=Unit())
    assert rel_fro(sm.read_kernel(Identity(), block_size=400).val, golden["n300_unit_K"]) < 1e-5
    dbx = SnpMemMap.write(str(tmp_path / "dbx.snp.memmap"), _bed("dbx"))                             # missing values in the file
    assert rel_fro(dbx.read_kernel(Unit(), block_size=30).val, golden["dbx_unit_K"]) < 1e-5


def test_intersect_apply_with_kernel(golden):
    from pysnptools_b200 import SnpKernel, Unit
    from pysnptools_b200.util import intersect_apply
    bed = _bed("n300")
    ids = bed.iid[np.arange(298, -1, -2)]                                     # every second individual, reversed
    kern, (vals, iid) = intersect_apply([SnpKernel(bed, Unit()), (np.arange(150.0).reshape(-1, 1), ids)])
    assert np.array_equal(kern.iid, bed.iid[::2]) and np.array_equal(iid, bed.iid[::2]) and vals[0, 0] == 149.0
    packed, n, m = fixture_packed("n300")
    from oracle import bed_oracle
    ref, _ = bed_oracle.read_kernel(packed, n, iid_index=np.arange(0, 300, 2))     # standardized on the 150 kept individuals
    assert rel_fro(kern.read().val, ref) < 1e-5
    late, _ = intersect_apply([SnpKernel(bed, Unit()), (vals, ids)], intersect_before_standardize=False)
    assert rel_fro(late.read().val, golden["n300_unit_K_every2"]) < 1e-5


def test_read_into_out_and_pinned_buffers(golden):
    """read(out=...) fills a caller-owned array (pinned memory from util.pinned_empty goes by direct DMA)."""
    from pysnptools_b200 import Unit
    from pysnptools_b200.util import pinned_empty
    bed = _bed("n300")
    want = i8_to_float(golden["n300_decode_i8"], np.float32)
    for order in ("F", "C"):
        buf = pinned_empty((300, 1015), dtype=np.float32, order=order)
        assert buf.flags["F_CONTIGUOUS" if order == "F" else "C_CONTIGUOUS"] and buf.flags["WRITEABLE"]
        d = bed.read(order=order, dtype=np.float32, out=buf)
        assert d.val is buf and np.array_equal(buf, want, equal_nan=True)
        d2 = bed.read(order=order, dtype=np.float32, standardizer=Unit(), out=buf)
        np.testing.assert_allclose(buf[:, :160], golden["n300_unit_val"], rtol=1e-5, atol=1e-6)
    sub = np.empty((150, 10), dtype=np.float64, order="F")
    assert bed[::2, :10].read(out=sub).val is sub and np.array_equal(sub, want[::2, :10].astype(np.float64), equal_nan=True)
    with pytest.raises(ValueError):
        bed.read(out=np.empty((3, 3)))
    del buf, d, d2
    import gc
    gc.collect()


def test_kernel_float64_exact_switch(golden):
    """set_kernel_float64("exact"): dtype=float64 kernels in float64 arithmetic on the GPU (file path, device-store path, in-memory values);
    float32 requests keep the tensor cores; the default ("tensor") is restored afterwards."""
    import pysnptools_b200 as p
    bed = p.Bed(os.path.join(DATA_DIR, "n300.bed"), count_A1=False)
    want = golden["n300_unit_K"]

    def rel(a, b):
        return np.linalg.norm(a - b) / np.linalg.norm(b)
    assert p.set_kernel_float64("exact") == "tensor"
    try:
        K = bed.read_kernel(p.Unit(), block_size=100).val                         # host path: pstb_snp_kernel_host_f64
        assert K.dtype == np.float64 and rel(K, want) < 1e-12
        sub = bed[::2, 5:900]
        x = sub.read().standardize(p.Unit()).val
        assert rel(sub.read_kernel(p.Unit()).val, x.dot(x.T)) < 1e-12
        sub.read(dtype=np.float32)                                                # puts the packed store on the device
        bed._store_for(None)
        assert rel(bed.read_kernel(p.Unit()).val, want) < 1e-12                   # device-store path: pstb_snp_kernel_f64
        data = bed.read()
        assert rel(data.read_kernel(p.Unit()).val, want) < 1e-12                  # in-memory values: pstb_float_kernel_f64
        K32 = bed.read_kernel(p.Unit(), dtype=np.float32).val
        assert K32.dtype == np.float32 and 1e-9 < rel(K32.astype(np.float64), want) < 1e-5
    finally:
        assert p.set_kernel_float64("tensor") == "exact"
    assert 1e-9 < rel(bed.read_kernel(p.Unit()).val, want) < 1e-5
    with pytest.raises(ValueError):
        p.set_kernel_float64("fp8")
