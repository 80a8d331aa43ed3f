"""CPU: pin the oracle (oracle/bed_oracle.py + oracle/c/pst_oracle.c) against the reference's golden vectors."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, SHAPES, fixture_packed, i8_to_float


@pytest.mark.parametrize("name", ["n300", "toydata", "dbx", "snpgen", "gen1", "gen4"])
def test_decode_matches_reference_fixture(name, golden, oracle):
    packed, n, m = fixture_packed(name)
    want = golden[name + "_decode_i8"]
    for dtype in (np.float64, np.float32, np.int8):
        for order in ("F", "C"):
            got = oracle.decode(packed, n, dtype=dtype, order=order)
            assert got.dtype == dtype and got.flags["F_CONTIGUOUS" if order == "F" else "C_CONTIGUOUS"]
            ref = want if dtype == np.int8 else i8_to_float(want, dtype)
            assert np.array_equal(got, ref, equal_nan=dtype != np.int8)
    if name + "_decode_A1_i8" in golden.files:
        got = oracle.decode(packed, n, count_A1=True, dtype=np.int8)
        assert np.array_equal(got, golden[name + "_decode_A1_i8"])


def test_decode_toydata_first10(golden, oracle):
    packed, n, m = fixture_packed("toydata")
    assert np.array_equal(oracle.decode(packed, n, sid_index=np.arange(10)), golden["toydata_decode_first10"])


def test_decode_subset_composed_by_reference_indexer(golden, oracle):
    packed, n, m = fixture_packed("n300")
    got = oracle.decode(packed, n, golden["n300_subset_rev_iid"], golden["n300_subset_rev_sid"], dtype=np.float32, order="C")
    assert np.array_equal(got, golden["n300_subset_rev_f32"], equal_nan=True)


@pytest.mark.parametrize("name", ["n300", "dbx", "snpgen"])
@pytest.mark.parametrize("spec", [("unit", False, np.nan, np.nan), ("beta_1_25", True, 1, 25), ("beta_2_10", True, 2, 10)])
def test_standardize_matches_reference_python_twin(name, spec, golden, oracle):
    tag, is_beta, a, b = spec
    packed, n, m = fixture_packed(name)
    raw = oracle.decode(packed, n)
    val, stats = oracle.standardize(raw, is_beta, a, b)
    want = golden["{0}_{1}_val".format(name, tag)]
    np.testing.assert_allclose(val[:, :want.shape[1]], want, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(stats, golden["{0}_{1}_stats".format(name, tag)], rtol=1e-12)


def test_trained_standardize(golden, oracle):
    packed, n, m = fixture_packed("n300")
    raw = oracle.decode(packed, n)
    _, st = oracle.standardize(raw[10:], False)
    np.testing.assert_allclose(st, golden["n300_trained_unit_stats"], rtol=1e-12)
    te, _ = oracle.standardize(raw[:10], False, use_stats=True, stats=st)
    np.testing.assert_allclose(te, golden["n300_trained_unit_test_val"], rtol=1e-12, atol=1e-13)
    _, stb = oracle.standardize(raw[10:], True, 1, 25)
    teb, _ = oracle.standardize(raw[:10], True, 1, 25, use_stats=True, stats=stb)
    np.testing.assert_allclose(teb, golden["n300_trained_beta_test_val"], rtol=1e-12, atol=1e-13)


def test_nan_and_snc_columns(golden, oracle):
    x = golden["n300_nancnc_input"]
    for tag, is_beta, a, b in (("unit", False, np.nan, np.nan), ("beta_1_25", True, 1, 25)):
        val, stats = oracle.standardize(x, is_beta, a, b)
        assert val[0, 0] == 0 and np.all(val[:, 1] == 0) and np.isinf(stats[1, 1])
        np.testing.assert_allclose(val, golden["n300_nancnc_{0}_val".format(tag)], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(stats, golden["n300_nancnc_{0}_stats".format(tag)], rtol=1e-12)


def test_kernel_goldens(golden, oracle):
    packed, n, m = fixture_packed("n300")
    K, _ = oracle.read_kernel(packed, n)
    assert "{0:.6f}".format(K[0, 0]) == "901.421836"
    np.testing.assert_allclose(K, golden["n300_unit_K"], rtol=1e-11, atol=1e-9)
    Kb, _ = oracle.read_kernel(packed, n, is_beta=True, a=1, b=25, block_size=500)
    np.testing.assert_allclose(Kb, golden["n300_beta_1_25_K"], rtol=1e-11, atol=1e-9)
    pk, nd, md = fixture_packed("dbx")
    Kd, _ = oracle.read_kernel(pk, nd, block_size=10)
    np.testing.assert_allclose(Kd, golden["dbx_unit_K"], rtol=1e-11, atol=1e-9)
    pt, nt, mt = fixture_packed("toydata")
    Kt, _ = oracle.read_kernel(pt, nt)
    ship = golden["toydata_unit_K_shipped"]
    assert np.linalg.norm(Kt - ship) / np.linalg.norm(ship) < 1e-13
    assert "{0:.6f}".format(Kt[0, 0] * nt / np.trace(Kt)) == "{0:.6f}".format(float(golden["toydata_unit_K_diagKtoN_00"]))
    np.testing.assert_allclose(K[::2, ::2], golden["n300_unit_K_every2"], rtol=1e-11, atol=1e-9)


def _c_oracle():
    so = os.path.join(ROOT, "oracle", "_build", "libpst_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return ctypes.CDLL(so)


def test_c_oracle_agrees_with_numpy_oracle(oracle):
    lib = _c_oracle()
    rng = np.random.default_rng(0)
    packed = oracle.synth_packed(203, 0, 37, missing_rate=0.1, seed=3)
    ii = rng.permutation(203)[:77].astype(np.int64)
    si = rng.permutation(37)[:19].astype(np.int64)
    p = ctypes.c_void_p
    for dtype, fn in ((np.float32, lib.pst_oracle_decode_f32), (np.float64, lib.pst_oracle_decode_f64), (np.int8, lib.pst_oracle_decode_i8)):
        for order_c in (0, 1):
            out = np.empty((77, 19), dtype=dtype, order="C" if order_c else "F")
            rc = fn(p(packed.ctypes.data), ctypes.c_int64(packed.shape[1]), ctypes.c_int64(203), ctypes.c_int64(37),
                    p(ii.ctypes.data), ctypes.c_int64(77), p(si.ctypes.data), ctypes.c_int64(19), 1, order_c, p(out.ctypes.data), 2)
            assert rc == 0
            assert np.array_equal(out, oracle.decode(packed, 203, ii, si, True, dtype), equal_nan=dtype != np.int8)
    raw = oracle.decode(packed, 203)
    for is_beta, a, b in ((0, np.nan, np.nan), (1, 1.0, 25.0)):
        want, wst = oracle.standardize(raw, bool(is_beta), a, b)
        for dtype, fn, tol in ((np.float64, lib.pst_oracle_standardize_f64, 1e-12), (np.float32, lib.pst_oracle_standardize_f32, 2e-6)):
            for order in ("F", "C"):
                val = np.array(raw, dtype=dtype, order=order)
                st = np.zeros((37, 2))
                fn(p(val.ctypes.data), ctypes.c_int64(203), ctypes.c_int64(37), int(order == "C"), is_beta, ctypes.c_double(a),
                   ctypes.c_double(b), 0, p(st.ctypes.data), 2)
                np.testing.assert_allclose(val, want, rtol=tol, atol=tol)
                np.testing.assert_allclose(st, wst, rtol=1e-12)


def test_synth_is_reproducible_by_range(oracle):
    a = oracle.synth_packed(50, 0, 20, 0.05, seed=1)
    b = oracle.synth_packed(50, 7, 13, 0.05, seed=1)
    assert np.array_equal(a[7:13], b)


def test_cross_kernel_oracle_pinned_by_trained_goldens(golden, oracle):
    """oracle.read_cross_kernel (the checker of pstb_snp_cross_kernel): train x test with the train statistics.  Its test-side values
    are the reference's own UnitTrained / BetaTrained outputs (goldens), and train x train degenerates to the symmetric kernel."""
    packed, n, m = fixture_packed("n300")
    train, test = np.arange(10, n), np.arange(10)
    for key, args in (("unit", {}), ("beta", dict(is_beta=True, a=1, b=25))):
        K, st = oracle.read_cross_kernel(packed, n, packed, n, iid_index_r=train, iid_index_c=test, **args)
        np.testing.assert_allclose(st, golden["n300_trained_{0}_stats".format(key)], rtol=1e-12)
        xr, _ = oracle.standardize(oracle.decode(packed, n, train), **args)
        np.testing.assert_allclose(K, xr @ golden["n300_trained_{0}_test_val".format(key)].T, rtol=1e-9, atol=1e-9)
    Ks, _ = oracle.read_cross_kernel(packed, n, packed, n)
    np.testing.assert_allclose(Ks, golden["n300_unit_K"], rtol=1e-10, atol=1e-8)
    # statistics given explicitly == statistics learned from the row side
    K2, _ = oracle.read_cross_kernel(packed, n, packed, n, iid_index_r=train, iid_index_c=test, stats=golden["n300_trained_unit_stats"])
    K1, _ = oracle.read_cross_kernel(packed, n, packed, n, iid_index_r=train, iid_index_c=test)
    np.testing.assert_allclose(K2, K1, rtol=1e-12, atol=1e-12)
