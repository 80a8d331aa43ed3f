"""GPU parity: K1 decode, K2 fused decode+standardize, K2f, gather, pack -- CUDA path vs the CPU oracle / goldens.

Integer / byte work is bit-exact (values and NaN pattern); standardized values are compared with the
reference's float64 path at rtol 1e-6 (north_star) -- in practice they agree to ~1e-15 (f64) / 1 ulp (f32).
"""
import ctypes

import numpy as np
import pytest

from conftest import SHAPES, fixture_packed, i8_to_float

pytestmark = pytest.mark.gpu

STD_RTOL = 1e-6


@pytest.fixture(scope="module")
def dev():
    import torch
    from pysnptools_b200 import device
    assert torch.cuda.is_available()
    return device


def _np(t):
    return t.cpu().numpy()


def _check_layout(t, order):
    assert (t.t().is_contiguous() if order == "F" else t.is_contiguous())


@pytest.mark.parametrize("name", list(SHAPES))
def test_decode_fixtures_bit_exact(name, golden, dev):
    packed, n, m = fixture_packed(name)
    store = dev.PackedStore.from_host(packed, n)
    for count_A1 in (False, True):
        key = name + ("_decode_A1_i8" if count_A1 else "_decode_i8")
        if key not in golden.files:
            continue
        want = golden[key]
        for dtype in (np.float32, np.float64, np.int8):
            for order in ("F", "C"):
                val, _ = dev.read(store, count_A1=count_A1, dtype=dtype, order=order)
                _check_layout(val, order)
                ref = want if dtype == np.int8 else i8_to_float(want, dtype)
                assert np.array_equal(_np(val), ref, equal_nan=dtype != np.int8), (name, count_A1, dtype, order)


def test_decode_golden_subset(golden, dev):
    packed, n, m = fixture_packed("n300")
    store = dev.PackedStore.from_host(packed, n)
    val, _ = dev.read(store, golden["n300_subset_rev_iid"], golden["n300_subset_rev_sid"], dtype=np.float32, order="C")
    assert np.array_equal(_np(val), golden["n300_subset_rev_f32"], equal_nan=True)


SHAPE_CASES = [(1, 3), (2, 5), (3, 1), (4, 4), (5, 9), (15, 7), (16, 33), (17, 2), (63, 40), (64, 65), (65, 3), (300, 70),
               (1000, 37), (4099, 21), (30001, 12), (500003, 3)]


@pytest.mark.parametrize("n,m", SHAPE_CASES)
def test_decode_random_shapes_and_selections(n, m, oracle, dev):
    rng = np.random.default_rng(n * 131 + m)
    packed = oracle.synth_packed(n, 0, m, missing_rate=0.1, seed=n)
    store = dev.PackedStore.from_host(packed, n)
    sels = [(None, None), (slice(None, None, -1), slice(None, None, -2)), (slice(n // 3, None), slice(1, None, 3)),
            (slice((n // 32) * 16, None), None),
            (rng.permutation(n)[: max(1, n // 2)], rng.permutation(m)[: max(1, m // 2)]),
            (rng.integers(0, n, size=min(n + 3, 50)), rng.integers(0, m, size=m + 2))]   # repeats allowed
    for isel, ssel in sels:
        for dtype, order, a1 in ((np.float32, "F", False), (np.float64, "F", True), (np.int8, "F", False), (np.float32, "C", True),
                                 (np.float64, "C", False), (np.int8, "C", True)):
            val, _ = dev.read(store, isel, ssel, count_A1=a1, dtype=dtype, order=order)
            ii = np.arange(n)[isel] if isinstance(isel, slice) else isel
            si = np.arange(m)[ssel] if isinstance(ssel, slice) else ssel
            ref = oracle.decode(packed, n, ii, si, a1, dtype, order)
            assert val.shape == ref.shape
            _check_layout(val, order)
            assert np.array_equal(_np(val), ref, equal_nan=dtype != np.int8), (n, m, dtype, order, a1)


def test_decode_unaligned_store_and_empty(oracle, dev):
    import torch
    packed = oracle.synth_packed(203, 0, 31, 0.2, seed=9)
    tight = torch.from_numpy(packed).cuda()                      # ld = 51: no 16-byte alignment -> non-bulk path
    store = dev.PackedStore(tight, 203, 31)
    assert store.ld == 51
    for order in ("F", "C"):
        val, _ = dev.read(store, dtype=np.float32, order=order)
        assert np.array_equal(_np(val), oracle.decode(packed, 203, dtype=np.float32), equal_nan=True)
        val, st = dev.read(store, dtype=np.float64, order=order, standardizer=("unit",))
        ref, rst = oracle.standardize(oracle.decode(packed, 203))
        np.testing.assert_allclose(_np(val), ref, rtol=1e-12, atol=1e-14)
    e, _ = dev.read(store, [], None, dtype=np.float64)
    assert tuple(e.shape) == (0, 31)
    e, _ = dev.read(store, None, [], dtype=np.float32, order="C")
    assert tuple(e.shape) == (203, 0)
    with pytest.raises(IndexError):
        dev.read(store, [203], None)
    with pytest.raises(IndexError):
        dev.read(store, None, np.array([31]))


@pytest.mark.parametrize("name", ["n300", "dbx", "snpgen"])
@pytest.mark.parametrize("spec", [("unit", ("unit",)), ("beta_1_25", ("beta", 1, 25)), ("beta_2_10", ("beta", 2, 10))])
def test_fused_standardize_vs_reference_goldens(name, spec, golden, dev):
    tag, std = spec
    packed, n, m = fixture_packed(name)
    store = dev.PackedStore.from_host(packed, n)
    want = golden["{0}_{1}_val".format(name, tag)]
    wst = golden["{0}_{1}_stats".format(name, tag)]
    for dtype in (np.float64, np.float32):
        for order in ("F", "C"):
            val, st = dev.read(store, dtype=dtype, order=order, standardizer=std)
            _check_layout(val, order)
            got = _np(val)[:, :want.shape[1]]
            assert not np.isnan(got).any()
            np.testing.assert_allclose(got, want, rtol=STD_RTOL, atol=1e-7 if dtype == np.float32 else 1e-12)
            if dtype == np.float64:
                np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-13)     # far inside the 1e-6 gate
            np.testing.assert_allclose(_np(st), wst, rtol=1e-12)


def test_trained_stats_and_snc(golden, dev):
    packed, n, m = fixture_packed("n300")
    store = dev.PackedStore.from_host(packed, n)
    _, st = dev.read(store, slice(10, None), None, standardizer=("unit",), dtype=np.float64)
    np.testing.assert_allclose(_np(st), golden["n300_trained_unit_stats"], rtol=1e-12)
    te, _ = dev.read(store, slice(0, 10), None, standardizer=("unit",), stats=st, dtype=np.float64)
    np.testing.assert_allclose(_np(te), golden["n300_trained_unit_test_val"], rtol=1e-11, atol=1e-13)
    assert abs(_np(te)[0, 0] - 0.23354968324845735) < 1e-14                      # standardizer.py:35-42 doctest
    _, stb = dev.read(store, slice(10, None), None, standardizer=("beta", 1, 25), dtype=np.float64)
    teb, _ = dev.read(store, slice(0, 10), None, standardizer=("beta", 1, 25), stats=stb, dtype=np.float64, order="C")
    np.testing.assert_allclose(_np(teb), golden["n300_trained_beta_test_val"], rtol=1e-11, atol=1e-13)
    # stats only (no output)
    none, st2 = dev.read(store, slice(10, None), None, standardizer=("unit",), want_out=False)
    assert none is None and np.array_equal(_np(st2), _np(st))


@pytest.mark.parametrize("n,m", [(5, 4), (203, 31), (4099, 21), (30001, 12), (500003, 3)])
def test_fused_standardize_random_vs_oracle(n, m, oracle, dev):
    rng = np.random.default_rng(n + m)
    packed = oracle.synth_packed(n, 0, m, missing_rate=0.07, seed=n + 1)
    packed[0, :] = 0xFF if n > 5 else packed[0, :]               # an SNC column (all dosage 2)
    store = dev.PackedStore.from_host(packed, n)
    isel = rng.permutation(n)[: max(2, n // 2)]
    for sel in (None, isel):
        raw = oracle.decode(packed, n, sel)
        for std, args in ((("unit",), (False, np.nan, np.nan)), (("beta", 1, 25), (True, 1, 25))):
            ref, rst = oracle.standardize(raw, *args)
            for dtype, order in ((np.float32, "F"), (np.float64, "C"), (np.float64, "F"), (np.float32, "C")):
                val, st = dev.read(store, sel, None, dtype=dtype, order=order, standardizer=std, count_A1=False)
                np.testing.assert_allclose(_np(val), ref, rtol=STD_RTOL, atol=1e-6 if dtype == np.float32 else 1e-12)
                np.testing.assert_allclose(_np(st), rst, rtol=1e-12)


def test_standardize_float_matrix_k2f(golden, oracle, dev):
    import torch
    x = golden["n300_nancnc_input"]
    for tag, std in (("unit", ("unit",)), ("beta_1_25", ("beta", 1, 25))):
        want, wst = golden["n300_nancnc_{0}_val".format(tag)], golden["n300_nancnc_{0}_stats".format(tag)]
        for dtype in (np.float64, np.float32):
            for order in ("C", "F"):
                t = torch.from_numpy(np.array(x, dtype=dtype, order="C")).cuda()
                if order == "F":
                    t = t.t().contiguous().t()
                st = dev.standardize(t, std)
                got = _np(t)
                assert got[0, 0] == 0 and np.all(got[:, 1] == 0) and np.isinf(_np(st)[1, 1])
                np.testing.assert_allclose(got, want, rtol=1e-12 if dtype == np.float64 else 1e-4, atol=1e-13 if dtype == np.float64 else 1e-6)
                np.testing.assert_allclose(_np(st), wst, rtol=1e-12 if dtype == np.float64 else 1e-6)
                # trained reuse (kernelreader/test.py:56-110)
                t2 = torch.from_numpy(np.array(x, dtype=dtype, order="C")).cuda()
                dev.standardize(t2, std, stats=st)
                np.testing.assert_allclose(_np(t2), got, rtol=1e-12, atol=1e-14)
    # non-genotype values take the general float path
    rng = np.random.default_rng(5)
    y = rng.normal(3.0, 2.0, size=(257, 19))
    y[rng.random(y.shape) < 0.1] = np.nan
    ref, rst = oracle.standardize(y)
    t = torch.from_numpy(y.copy()).cuda()
    st = dev.standardize(t, ("unit",))
    np.testing.assert_allclose(_np(t), ref, rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(_np(st), rst, rtol=1e-12)


def test_pack_round_trip_and_illegal_values(oracle, dev):
    import torch
    for n, m in ((1, 1), (2, 3), (5, 4), (190, 20), (1000, 17)):
        packed = oracle.synth_packed(n, 0, m, 0.15, seed=n)
        for a1 in (False, True):
            for dtype in (np.float32, np.float64, np.int8):
                raw = oracle.decode(packed, n, count_A1=a1, dtype=dtype, order="F")
                for order in ("F", "C"):
                    t = torch.from_numpy(np.ascontiguousarray(raw)).cuda()
                    if order == "F":
                        t = t.t().contiguous().t()
                    store = dev.pack(t, count_A1=a1)
                    got = _np(store.tensor)[:, : (n + 3) // 4]
                    assert np.array_equal(got, packed), (n, m, a1, dtype, order)
    bad = torch.full((5, 3), 5.0, device="cuda")
    with pytest.raises(ValueError):
        dev.pack(bad)


def _host_read(lib, packed, n, m, ii, si, dtype, order, mode=0, a=0.0, b=0.0, stats=None, count_a1=0, out=None):
    code = {np.float32: 0, np.float64: 1, np.int8: 2}[dtype]
    ni = n if ii is None else len(ii)
    ns = m if si is None else len(si)
    if out is None:
        out = np.full((ni, ns), 7, dtype=dtype, order=order)
    st = np.zeros((ns, 2)) if stats is None else np.array(stats, dtype=np.float64)
    p = ctypes.c_void_p
    rc = lib.pstb_read_host(p(packed.ctypes.data), n, m, p(ii.ctypes.data) if ii is not None else None, ni,
                            p(si.ctypes.data) if si is not None else None, ns, count_a1, mode, a, b, int(stats is not None),
                            p(st.ctypes.data), p(out.ctypes.data), code, 0 if order == "F" else 1)
    return rc, out, st


def test_host_abi_read(oracle):
    from pysnptools_b200 import _lib
    lib = _lib.lib
    rng = np.random.default_rng(2)
    n, m = 1003, 257
    packed = oracle.synth_packed(n, 0, m, 0.05, seed=4)
    ii = rng.permutation(n)[:400].astype(np.int64)
    si = rng.permutation(m)[:100].astype(np.int64)
    si_run = np.arange(20, 120, dtype=np.int64)
    for isel, ssel in ((None, None), (ii, si), (None, si_run), (ii[::-1].copy(), None)):
        for dtype in (np.float32, np.float64, np.int8):
            for order in ("F", "C"):
                rc, out, _ = _host_read(lib, packed, n, m, isel, ssel, dtype, order)
                assert rc == 0, _lib.last_error()
                assert np.array_equal(out, oracle.decode(packed, n, isel, ssel, False, dtype, order), equal_nan=dtype != np.int8)
        raw = oracle.decode(packed, n, isel, ssel)
        ref, rst = oracle.standardize(raw, True, 1, 25)
        rc, out, st = _host_read(lib, packed, n, m, isel, ssel, np.float64, "F", mode=2, a=1.0, b=25.0)
        assert rc == 0, _lib.last_error()
        np.testing.assert_allclose(out, ref, rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(st, rst, rtol=1e-12)
        rc, out2, _ = _host_read(lib, packed, n, m, isel, ssel, np.float32, "C", mode=2, a=1.0, b=25.0, stats=st)
        assert rc == 0
        np.testing.assert_allclose(out2, ref, rtol=STD_RTOL, atol=1e-6)
    bad = np.array([m], dtype=np.int64)
    rc, _, _ = _host_read(lib, packed, n, m, None, bad, np.float32, "F")
    assert rc != 0 and "out of range" in _lib.last_error()


def test_host_abi_multichunk_pinned_and_pageable(oracle):
    """> 64 MiB of output forces several chunks through both streams; pinned and pageable destinations."""
    from pysnptools_b200 import _lib
    lib = _lib.lib
    n, m = 4000, 6001
    packed = np.tile(oracle.synth_packed(n, 0, 400, 0.03, seed=8), (16, 1))[:m]
    ref = oracle.decode(packed, n, dtype=np.float32)
    refs, rst = oracle.standardize(ref[:, :500])
    for order in ("F", "C"):
        rc, out, st = _host_read(lib, packed, n, m, None, None, np.float32, order, mode=1)
        assert rc == 0, _lib.last_error()
        np.testing.assert_allclose(out[:, :500], refs, rtol=STD_RTOL, atol=1e-6)
        np.testing.assert_allclose(st[:500], rst, rtol=1e-12)
        assert np.array_equal(out[:, 400:800], out[:, 4000:4400])               # tiled input -> periodic output
    nbytes = n * m * 4
    ptr = lib.pstb_host_alloc(nbytes)
    assert ptr
    try:
        buf = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)), shape=(n * m,))
        for order in ("F", "C"):
            out = buf.reshape((n, m), order=order)
            rc, out, _ = _host_read(lib, packed, n, m, None, None, np.float32, order, out=out)
            assert rc == 0, _lib.last_error()
            assert np.array_equal(out, ref, equal_nan=True)
    finally:
        lib.pstb_host_free(ptr)


def test_host_abi_standardize_and_subset(golden, oracle):
    from pysnptools_b200 import _lib
    lib = _lib.lib
    p = ctypes.c_void_p
    x = golden["n300_nancnc_input"]
    for dtype, code in ((np.float64, 1), (np.float32, 0)):
        for order in ("F", "C"):
            val = np.array(x, dtype=dtype, order=order)
            st = np.zeros((x.shape[1], 2))
            rc = lib.pstb_standardize_host(p(val.ctypes.data), code, 0 if order == "F" else 1, x.shape[0], x.shape[1], 1,
                                           float("nan"), float("nan"), 1, 0, p(st.ctypes.data))
            assert rc == 0, _lib.last_error()
            np.testing.assert_allclose(val, golden["n300_nancnc_unit_val"], rtol=1e-12 if dtype == np.float64 else 1e-4, atol=1e-6)
    rng = np.random.default_rng(0)
    for v in (1, 3):
        src = rng.normal(size=(23, 17, v))
        rows = np.array([5, 0, 22, 5], dtype=np.int64)
        cols = np.arange(16, -1, -2, dtype=np.int64)
        for oi in ("C", "F"):
            for oo in ("C", "F"):
                for dti, dto, ci, co in ((np.float64, np.float64, 1, 1), (np.float32, np.float64, 0, 1), (np.float32, np.float32, 0, 0)):
                    a = np.array(src, dtype=dti, order=oi)
                    out = np.zeros((4, 9, v), dtype=dto, order=oo)
                    rc = lib.pstb_subset_host(p(a.ctypes.data), ci, 0 if oi == "F" else 1, 23, 17, v, p(rows.ctypes.data), 4,
                                              p(cols.ctypes.data), 9, p(out.ctypes.data), co, 0 if oo == "F" else 1)
                    assert rc == 0, _lib.last_error()
                    assert np.array_equal(out, a[rows][:, cols].astype(dto))


def test_launch_counter_moves(dev, oracle):
    from pysnptools_b200 import _lib
    before = _lib.lib.pstb_launch_count()
    packed = oracle.synth_packed(64, 0, 8, 0.0, seed=0)
    dev.read(dev.PackedStore.from_host(packed, 64), dtype=np.float32)
    assert _lib.lib.pstb_launch_count() > before


@pytest.mark.parametrize("n,m", [(203, 31), (4099, 21), (70001, 9)])
def test_standardize_with_repeated_and_reversed_indices(n, m, oracle, dev):
    """Statistics are over the SELECTED individuals with multiplicity: a repeated index must count twice
    (the selection-mask shortcut of the gather kernel has to fall back)."""
    rng = np.random.default_rng(n)
    packed = oracle.synth_packed(n, 0, m, missing_rate=0.1, seed=n + 7)
    store = dev.PackedStore.from_host(packed, n)
    for ii in (rng.integers(0, n, size=n // 2 + 3), np.arange(n)[::-3], np.concatenate([rng.permutation(n)[:50], [0, 0, n - 1]])):
        raw = oracle.decode(packed, n, ii)
        for std, args in ((("unit",), (False, np.nan, np.nan)), (("beta", 2, 10), (True, 2, 10))):
            ref, rst = oracle.standardize(raw, *args)
            for dtype, order in ((np.float64, "F"), (np.float32, "F"), (np.float64, "C")):
                val, st = dev.read(store, ii, None, dtype=dtype, order=order, standardizer=std)
                np.testing.assert_allclose(_np(st), rst, rtol=1e-12)
                np.testing.assert_allclose(_np(val), ref, rtol=STD_RTOL, atol=1e-6 if dtype == np.float32 else 1e-12)


def test_all_missing_and_constant_snps(oracle, dev):
    """Unpinned edges, python-twin behaviour: an all-missing SNP gives zeros with NaN statistics; an SNC SNP gives sd = inf."""
    n, m = 77, 6
    packed = oracle.synth_packed(n, 0, m, 0.1, seed=1)
    packed[2, :] = 0x55                                            # every code 01 = missing
    packed[4, :] = 0x00                                            # every dosage 0: no variation
    store = dev.PackedStore.from_host(packed, n)
    for std in (("unit",), ("beta", 1, 25)):
        for order in ("F", "C"):
            val, st = dev.read(store, dtype=np.float64, order=order, standardizer=std)
            v, s = _np(val), _np(st)
            assert np.all(v[:, 2] == 0) and np.isnan(s[2]).all()
            assert np.all(v[:, 4] == 0) and s[4, 0] == 0 and np.isinf(s[4, 1])
            assert not np.isnan(v).any()
    K, _ = dev.snp_kernel(store, chunk=64)
    x, _ = oracle.standardize(oracle.decode(packed[[0, 1, 3, 5]], n))
    ref = x @ x.T
    assert np.linalg.norm(_np(K).astype(np.float64) - ref) / np.linalg.norm(ref) < 1e-5


@pytest.mark.parametrize("n,m", [(515, 264), (1030, 520), (37, 1032), (2049, 256), (600, 776)])
def test_c_order_wide_tiles(n, m, oracle, dev):
    """C order with 32-byte row pieces per lane (k_emit_c_wide: float32 / float64, n_sid a multiple of 8 / 4 and >= one tile):
    partial SNP tiles, partial row quads, gathered / strided axes, count_A1, fused standardize, and the narrow fallback."""
    rng = np.random.default_rng(n + m)
    packed = oracle.synth_packed(n, 0, m, missing_rate=0.1, seed=n * m)
    store = dev.PackedStore.from_host(packed, n)
    sels = [(None, None), (rng.permutation(n)[: max(1, n // 2)], None), (slice(None, None, -1), slice(0, m - m % 8 - 8)),
            (None, rng.permutation(m)[: 264]), (slice(16, None), slice(3, 3 + 256))]
    for isel, ssel in sels:
        ii = None if isel is None else np.arange(n)[isel]
        si = None if ssel is None else np.arange(m)[ssel]
        for dtype in (np.float32, np.float64):
            for a1 in (False, True):
                val, _ = dev.read(store, isel, ssel, count_A1=a1, dtype=dtype, order="C")
                assert val.is_contiguous()
                assert np.array_equal(_np(val), oracle.decode(packed, n, ii, si, a1, dtype, "C"), equal_nan=True), (n, m, dtype, a1)
            for std, args in ((("unit",), {}), (("beta", 1, 25), dict(is_beta=True, a=1, b=25))):
                val, st = dev.read(store, isel, ssel, dtype=dtype, order="C", standardizer=std)
                ref, rst = oracle.standardize(oracle.decode(packed, n, ii, si), **args)
                np.testing.assert_allclose(_np(val), ref, rtol=1e-6 if dtype == np.float32 else 1e-12, atol=1e-6 if dtype == np.float32 else 1e-12)
                np.testing.assert_allclose(st.cpu().numpy(), rst, rtol=1e-12)
                valf, _ = dev.read(store, isel, ssel, dtype=dtype, order="F", standardizer=std)
                assert np.array_equal(_np(val), _np(valf))                       # C and F order carry identical values


@pytest.mark.parametrize("n,m", [(1024, 40), (4100, 333), (30000, 20), (52000, 7), (2051, 9)])
def test_standardize_float_matrix_staged_columns(n, m, oracle, dev):
    """K2f, F order with the column staged in shared memory (TMA bulk load + bulk store; one or two buffers; (2051, 9): float32
    columns that are not 16-byte multiples fall back to the sweeping kernel): values, statistics, trained reuse, stats only."""
    import torch
    from pysnptools_b200 import _lib
    rng = np.random.default_rng(n)
    y = rng.integers(0, 3, size=(n, m)).astype(np.float64)
    y[:, 1] = 1.0                                                       # SNC column
    y[rng.random(y.shape) < 0.07] = np.nan
    if m > 5:
        y[:, 5] = rng.normal(50.0, 0.01, size=n)                        # |mean| >> sd: the two-pass formula matters
    for std, args in ((("unit",), {}), (("beta", 1, 25), dict(is_beta=True, a=1, b=25))):
        src = y if std[0] == "unit" else np.where(np.abs(y) > 2, np.nan, y)
        # column 5 (mean 50, sd 0.01) amplifies a 4-ulp difference in the mean (summation order) by mean / sd = 5000
        for dtype, rtol, atol in ((np.float64, 1e-10, 1e-10), (np.float32, 2e-4, 2e-5)):
            ref, rst = oracle.standardize(np.array(src, dtype=dtype).astype(np.float64), **args)     # the values the GPU is given
            # C order first: one statistics sweep (shifted sums per row block, merged pairwise) + one transform sweep
            tc = torch.from_numpy(np.array(src, dtype=dtype, order="C")).cuda()
            stc = dev.standardize(tc, std)
            np.testing.assert_allclose(_np(tc), ref, rtol=rtol, atol=atol)
            finc = np.isfinite(rst).all(axis=1)
            np.testing.assert_allclose(_np(stc)[finc], rst[finc], rtol=1e-12 if dtype == np.float64 else 1e-6)
            assert np.array_equal(np.isinf(_np(stc)[:, 1]), np.isinf(rst[:, 1]))
            t = torch.from_numpy(np.array(src, dtype=dtype, order="F")).cuda()
            assert t.t().is_contiguous()
            st = dev.standardize(t, std)
            got, gst = _np(t), _np(st)
            fin = np.isfinite(rst).all(axis=1)
            np.testing.assert_allclose(gst[fin], rst[fin], rtol=1e-12 if dtype == np.float64 else 1e-6)
            assert np.array_equal(np.isinf(gst[:, 1]), np.isinf(rst[:, 1]))
            np.testing.assert_allclose(got, ref, rtol=rtol, atol=atol)
            t2 = torch.from_numpy(np.array(src, dtype=dtype, order="F")).cuda()
            dev.standardize(t2, std, stats=st)
            np.testing.assert_allclose(_np(t2), got, rtol=1e-12, atol=1e-14)
            t3 = torch.from_numpy(np.array(src, dtype=dtype, order="F")).cuda()
            st3 = dev.standardize(t3, std, apply_in_place=False)        # statistics only: the matrix is left alone
            assert np.array_equal(_np(t3), np.array(src, dtype=dtype), equal_nan=True) and np.array_equal(_np(st3), gst, equal_nan=True)


@pytest.mark.parametrize("pinned", [False, True])
def test_host_abi_standardize_pipelined_chunks(pinned, oracle, monkeypatch):
    """pstb_standardize_host, F order: many column blocks through the 4-slot H2D / kernel / D2H ring (pageable arrays are staged
    by host threads, pinned ones copied directly); statistics only (apply_in_place = 0) leaves the array alone; trained reuse."""
    from pysnptools_b200 import _lib
    from pysnptools_b200.util import pinned_empty
    lib = _lib.lib
    p = ctypes.c_void_p
    monkeypatch.setenv("PSTB_STD_HOST_CHUNK_KB", "96")
    rng = np.random.default_rng(3)
    n, m = 1500, 173
    y = rng.integers(0, 3, size=(n, m)).astype(np.float64)
    y[rng.random(y.shape) < 0.05] = np.nan
    for dtype, code, rtol, atol in ((np.float64, 1, 1e-12, 1e-13), (np.float32, 0, 1e-4, 1e-6)):
        for mode, args, ab in ((1, {}, (float("nan"), float("nan"))), (2, dict(is_beta=True, a=1, b=25), (1.0, 25.0))):
            ref, rst = oracle.standardize(y, **args)
            val = pinned_empty((n, m), dtype=dtype, order="F") if pinned else np.empty((n, m), dtype=dtype, order="F")
            val[...] = y
            st = np.zeros((m, 2))
            assert lib.pstb_standardize_host(p(val.ctypes.data), code, 0, n, m, mode, ab[0], ab[1], 0, 0, p(st.ctypes.data)) == 0, _lib.last_error()
            assert np.array_equal(val, y.astype(dtype), equal_nan=True)                       # statistics only
            np.testing.assert_allclose(st, rst, rtol=1e-12 if dtype == np.float64 else 1e-6)
            assert lib.pstb_standardize_host(p(val.ctypes.data), code, 0, n, m, mode, ab[0], ab[1], 1, 0, p(st.ctypes.data)) == 0, _lib.last_error()
            np.testing.assert_allclose(val, ref, rtol=rtol, atol=atol)
            again = np.array(y, dtype=dtype, order="F")
            assert lib.pstb_standardize_host(p(again.ctypes.data), code, 0, n, m, mode, ab[0], ab[1], 1, 1, p(st.ctypes.data)) == 0, _lib.last_error()
            np.testing.assert_allclose(again, val, rtol=1e-12 if dtype == np.float64 else 1e-6, atol=1e-7)


@pytest.mark.parametrize("n,m", [(256, 37), (1000, 131), (4104, 67), (10000, 94), (12296, 23), (24568, 11)])
def test_dynamic_record_feed(n, m, oracle, dev):
    """The F-order kernels hand their records out through an atomic counter (one claim ahead of the record being written): every
    output column must still be written exactly once whatever the claim order.  Dense rows over the warp-per-record sizes, SNP
    counts that are not multiples of anything, scattered / strided / reversed SNP selections, a row range starting at a multiple
    of 16, trained statistics, count_A1 -- decode bit-exact, standardized values at the reference tolerance."""
    from pysnptools_b200 import _lib
    rng = np.random.default_rng(n + m)
    packed = oracle.synth_packed(n, 0, m, missing_rate=0.04, seed=n)
    store = dev.PackedStore.from_host(packed, n)
    sels = [None, np.arange(m)[::-1].copy(), rng.permutation(m)[: m // 2 + 1].astype(np.int64), np.arange(1, m, 3, dtype=np.int64), np.array([5], dtype=np.int64)]
    rows = [None]
    if n >= 1024:
        rows.append(np.arange(16, 16 + (n - 16) // 8 * 8 - 8, dtype=np.int64))            # dense range, length a multiple of 8
    before = _lib.lib.pstb_launch_count()
    for ii in rows:
        for si in sels:
            for count_A1 in (False, True):
                want = oracle.decode(packed, n, ii, si, count_A1, np.float64, "F")
                for dtype in (np.float32, np.float64):
                    val, _ = dev.read(store, ii, si, count_A1=count_A1, dtype=dtype, order="F")
                    assert np.array_equal(_np(val), want.astype(dtype), equal_nan=True), (n, m, dtype, count_A1)
            raw = oracle.decode(packed, n, ii, si)
            for std, args in ((("unit",), (False, np.nan, np.nan)), (("beta", 1, 25), (True, 1, 25))):
                ref, rst = oracle.standardize(raw, *args)
                for dtype in (np.float64, np.float32):
                    val, st = dev.read(store, ii, si, dtype=dtype, order="F", standardizer=std)
                    np.testing.assert_allclose(_np(st), rst, rtol=1e-12)
                    np.testing.assert_allclose(_np(val), ref, rtol=STD_RTOL, atol=1e-6 if dtype == np.float32 else 1e-12)
                    val2, _ = dev.read(store, ii, si, dtype=dtype, order="F", standardizer=std, stats=st)
                    assert np.array_equal(_np(val2), _np(val))                              # trained statistics: the same table, the same bits
    assert _lib.lib.pstb_launch_count() > before


@pytest.mark.parametrize("n,m,gather", [(4104, 6000, False), (10000, 2600, False), (300, 40000, False), (30001, 700, False), (10000, 1500, True),
                                        (2049, 3000, True)])
def test_dynamic_record_feed_more_records_than_groups(n, m, gather, oracle, dev):
    """Several claims per warp / CTA: more SNP records than the launch has resident groups (1 184 warps, 296 or 148 CTAs), on the
    warp-per-record (8 x 1 and 4 x 2 CTAs per SM), short-record, CTA-per-record and gathered (batches of four) kernels, plus the
    statistics-only pass behind a C-order read -- every column written once, from the right record, with the right statistics."""
    packed = oracle.synth_packed(n, 0, m, missing_rate=0.03, seed=n + m)
    store = dev.PackedStore.from_host(packed, n)
    ii = np.random.default_rng(m).permutation(n)[: n // 2].astype(np.int64) if gather else None
    raw = oracle.decode(packed, n, ii, None)
    for dtype in (np.float32, np.int8) if not gather else (np.float32,):
        val, _ = dev.read(store, ii, None, dtype=dtype, order="F")
        want = oracle.decode(packed, n, ii, None, False, dtype, "F")
        assert np.array_equal(_np(val), want, equal_nan=dtype != np.int8), (n, m, dtype)
    ref, rst = oracle.standardize(raw, False, np.nan, np.nan)
    for order in ("F", "C"):
        val, st = dev.read(store, ii, None, dtype=np.float32, order=order, standardizer=("unit",))
        np.testing.assert_allclose(_np(st), rst, rtol=1e-12)
        np.testing.assert_allclose(_np(val), ref, rtol=STD_RTOL, atol=1e-6)
    # K2f on the decoded matrix (k_std_f_staged hands its columns out the same way)
    import torch
    x = torch.from_numpy(np.ascontiguousarray(raw.astype(np.float32).T)).cuda().t()          # [n, m], F-contiguous on the device
    st2 = dev.standardize(x, ("unit",))
    fin = np.isfinite(rst).all(axis=1)
    np.testing.assert_allclose(_np(st2)[fin], rst[fin], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(_np(x)[:, fin], ref[:, fin], rtol=1e-5, atol=1e-5)
