"""GPU parity: K3 kinship K = X X^T on tcgen05 vs the float64 oracle / the reference's golden kernels.

Gate (north_star): relative Frobenius error <= 1e-5.  The fp16 hi/lo split lands near 1e-7.
"""
import os

import numpy as np
import pytest

from conftest import fixture_packed

pytestmark = pytest.mark.gpu
K_TOL = 1e-5


def rel_fro(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.fixture(scope="module")
def dev():
    import torch
    from pysnptools_b200 import device
    assert torch.cuda.is_available()
    return device


@pytest.mark.parametrize("n,k", [(100, 64), (300, 128), (640, 448), (1000, 64)])
def test_syrk_planes_tensor_core_stage(n, k, dev):
    """The tcgen05 stage alone: K_lower = hi hi^T + hi lo^T + lo hi^T on given fp16 planes."""
    import torch
    from pysnptools_b200._lib import lib, check
    g = torch.Generator(device="cuda").manual_seed(n + k)
    n_pad = (n + 255) // 256 * 256
    x = torch.randn((n_pad, k), generator=g, device="cuda", dtype=torch.float32) * 3.0
    x[n:] = 0
    hi = x.to(torch.float16)
    lo = (x - hi.float()).to(torch.float16)
    K = torch.full((n, n), 7.0, device="cuda", dtype=torch.float32)
    check(lib.pstb_syrk_planes(hi.data_ptr(), lo.data_ptr(), n, n_pad, k, K.data_ptr(), n, 0, 0.5, torch.cuda.current_stream().cuda_stream))
    h64, l64 = hi[:n].double(), lo[:n].double()
    ref = 0.5 * (h64 @ h64.T + h64 @ l64.T + l64 @ h64.T)
    got = torch.tril(K).double().cpu().numpy()
    want = torch.tril(ref).cpu().numpy()
    assert rel_fro(got, want) < 3e-6, rel_fro(got, want)
    # accumulate adds onto the existing lower triangle
    check(lib.pstb_syrk_planes(hi.data_ptr(), lo.data_ptr(), n, n_pad, k, K.data_ptr(), n, 1, 0.5, torch.cuda.current_stream().cuda_stream))
    assert rel_fro(torch.tril(K).double().cpu().numpy(), 2 * want) < 3e-6
    check(lib.pstb_mirror_lower(K.data_ptr(), n, n, torch.cuda.current_stream().cuda_stream))
    Kc = K.cpu().numpy()
    assert np.array_equal(Kc, Kc.T)


def test_kernel_goldens(golden, dev):
    packed, n, m = fixture_packed("n300")
    store = dev.PackedStore.from_host(packed, n)
    K, st = dev.snp_kernel(store)
    Kc = K.double().cpu().numpy()
    assert rel_fro(Kc, golden["n300_unit_K"]) < K_TOL
    assert abs(Kc[0, 0] - 901.421836) < 901.421836 * K_TOL                         # snpreader.py:308-313 doctest
    np.testing.assert_allclose(st.cpu().numpy(), golden["n300_unit_stats"], rtol=1e-12)
    Kb, _ = dev.snp_kernel(store, standardizer=("beta", 1, 25), chunk=512)
    assert rel_fro(Kb.double().cpu().numpy(), golden["n300_beta_1_25_K"]) < K_TOL
    pd, nd, md = fixture_packed("dbx")
    Kd, _ = dev.snp_kernel(dev.PackedStore.from_host(pd, nd), chunk=64)            # 2 chunks, missing values
    assert rel_fro(Kd.double().cpu().numpy(), golden["dbx_unit_K"]) < K_TOL
    pt, nt, mt = fixture_packed("toydata")
    Kt, _ = dev.snp_kernel(dev.PackedStore.from_host(pt, nt))
    Ktc = Kt.double().cpu().numpy()
    assert rel_fro(Ktc, golden["toydata_unit_K_shipped"]) < K_TOL
    assert abs(np.trace(Ktc) - 5_000_000) < 50                                       # trace(K) = N * M for Unit
    assert abs(Ktc[0, 0] * nt / np.trace(Ktc) - float(golden["toydata_unit_K_diagKtoN_00"])) < 1e-5


@pytest.mark.parametrize("n,m,chunk", [(257, 300, 64), (1500, 2000, 1024), (2100, 700, None)])
def test_kernel_random_vs_oracle(n, m, chunk, oracle, dev):
    rng = np.random.default_rng(n)
    packed = oracle.synth_packed(n, 0, m, missing_rate=0.05, seed=n)
    store = dev.PackedStore.from_host(packed, n)
    for std, args in ((("unit",), {}), (("beta", 1, 25), dict(is_beta=True, a=1, b=25))):
        ref, rst = oracle.read_kernel(packed, n, **args)
        K, st = dev.snp_kernel(store, standardizer=std, chunk=chunk)
        Kc = K.double().cpu().numpy()
        assert np.array_equal(Kc, Kc.T)
        assert rel_fro(Kc, ref) < K_TOL, rel_fro(Kc, ref)
        np.testing.assert_allclose(st.cpu().numpy(), rst, rtol=1e-12)
    # iid gather + sid subset + count_A1 + trained statistics
    ii = rng.permutation(n)[: n // 2]
    si = np.sort(rng.permutation(m)[: m // 2])
    ref, rst = oracle.read_kernel(packed, n, iid_index=ii, sid_index=si, count_A1=True)
    K, st = dev.snp_kernel(store, ii, si, count_A1=True, chunk=chunk)
    assert rel_fro(K.double().cpu().numpy(), ref) < K_TOL
    K2, _ = dev.snp_kernel(store, ii, si, count_A1=True, stats=st, chunk=chunk)
    assert rel_fro(K2.double().cpu().numpy(), ref) < K_TOL
    # SNP-sharded accumulation == one pass (what the multi-GPU path sums with NCCL)
    half = len(si) // 2
    Ka, _ = dev.snp_kernel(store, ii, si[:half], count_A1=True, chunk=chunk, mirror=False)
    Ka, _ = dev.snp_kernel(store, ii, si[half:], count_A1=True, chunk=chunk, K=Ka, accumulate=True)
    assert rel_fro(Ka.double().cpu().numpy(), ref) < K_TOL


def test_convert_kernel(dev):
    import torch
    K = torch.randn((37, 37), device="cuda")
    out = dev.convert_kernel(K, np.float64, scale=2.0)
    assert out.dtype == torch.float64 and torch.allclose(out, K.double() * 2.0)


@pytest.mark.parametrize("n,m,world", [(700, 300, 1), (1500, 700, 3), (2100, 400, 8), (200, 130, 3)])   # last: more ranks than tiles
def test_kernel_tile_sharding(n, m, world, oracle, dev):
    """K-tile sharding (cfg5 path): every rank's tiles together are exactly K; no tile is owned twice."""
    packed = oracle.synth_packed(n, 0, m, missing_rate=0.03, seed=n + world)
    store = dev.PackedStore.from_host(packed, n)
    ref, rst = oracle.read_kernel(packed, n)
    K = np.full((n, n), np.nan)
    seen = set()
    T = (n + 255) // 256
    for rank in range(world):
        tiles, coords, st = dev.snp_kernel_tiles(store, rank=rank, world=world, chunk=256)   # two chunks: accumulate path
        np.testing.assert_allclose(st.cpu().numpy(), rst, rtol=1e-12)
        th = tiles.double().cpu().numpy()
        for t, (I, J) in enumerate(coords):
            assert J <= I and (I, J) not in seen
            seen.add((int(I), int(J)))
            r0, c0 = I * 256, J * 256
            r1, c1 = min(n, r0 + 256), min(n, c0 + 256)
            K[r0:r1, c0:c1] = th[t, : r1 - r0, : c1 - c0]
            if I != J:
                K[c0:c1, r0:r1] = th[t, : r1 - r0, : c1 - c0].T
    assert len(seen) == T * (T + 1) // 2 and not np.isnan(K).any()
    assert np.linalg.norm(K - ref) / np.linalg.norm(ref) < K_TOL
    assert np.allclose(K, K.T, rtol=0, atol=1e-3 * np.abs(ref).max())      # diagonal tiles are stored whole (both triangles)


@pytest.mark.parametrize("n_r,n_c,m,chunk", [(300, 100, 200, 64), (1500, 700, 1000, 512), (257, 1030, 333, None), (5, 2100, 64, None)])
def test_cross_kernel_vs_oracle(n_r, n_c, m, chunk, oracle, dev):
    """Train x test kernel (pstb_snp_cross_kernel): both sides standardized with the train statistics, missing -> 0."""
    rng = np.random.default_rng(n_r + n_c)
    pr = oracle.synth_packed(n_r, 0, m, missing_rate=0.04, seed=n_r)
    pc = oracle.synth_packed(n_c, 0, m, missing_rate=0.04, seed=n_c + 1)
    sr, sc = dev.PackedStore.from_host(pr, n_r), dev.PackedStore.from_host(pc, n_c)
    for std, args in ((("unit",), {}), (("beta", 1, 25), dict(is_beta=True, a=1, b=25))):
        ref, rst = oracle.read_cross_kernel(pr, n_r, pc, n_c, **args)
        out, st = dev.snp_cross_kernel(sr, sc, standardizer=std, chunk=chunk)
        assert tuple(out.shape) == (n_r, n_c)
        assert rel_fro(out.double().cpu().numpy(), ref) < K_TOL, rel_fro(out.double().cpu().numpy(), ref)
        np.testing.assert_allclose(st.cpu().numpy(), rst, rtol=1e-12)
        out2, _ = dev.snp_cross_kernel(sr, sc, standardizer=std, stats=st, chunk=chunk)       # given (trained) statistics
        assert rel_fro(out2.double().cpu().numpy(), ref) < K_TOL
    # gathered iids on both sides, differently ordered SNP selections pairing the same SNPs, count_A1 on the test side only
    ii_r, ii_c = rng.permutation(n_r)[: max(1, n_r // 2)], rng.permutation(n_c)[: max(1, n_c // 3)]
    si = rng.permutation(m)[: m // 2]
    ref, _ = oracle.read_cross_kernel(pr, n_r, pc, n_c, iid_index_r=ii_r, iid_index_c=ii_c, sid_index_r=si, sid_index_c=si)
    out, st = dev.snp_cross_kernel(sr, sc, ii_r, ii_c, si, si, chunk=chunk)
    assert rel_fro(out.double().cpu().numpy(), ref) < K_TOL
    # SNP-sharded accumulation (what the multi-GPU path sums) == one pass
    half = len(si) // 2
    acc, st_a = dev.snp_cross_kernel(sr, sc, ii_r, ii_c, si[:half], si[:half], chunk=chunk)
    acc, st_b = dev.snp_cross_kernel(sr, sc, ii_r, ii_c, si[half:], si[half:], chunk=chunk, out=acc, accumulate=True)
    assert rel_fro(acc.double().cpu().numpy(), ref) < K_TOL
    # train x train through the rectangular path == the symmetric kernel
    Ks, _ = dev.snp_kernel(sr, chunk=chunk)
    Kx, _ = dev.snp_cross_kernel(sr, sr, chunk=chunk)
    assert rel_fro(Kx.double().cpu().numpy(), Ks.double().cpu().numpy()) < 4e-6      # L_i B_k above the diagonal here, the mirrored L_k B_i there


def test_cross_kernel_edges(oracle, dev):
    import torch
    pr = oracle.synth_packed(40, 0, 70, seed=3)
    s = dev.PackedStore.from_host(pr, 40)
    out, st = dev.snp_cross_kernel(s, s, [1, 2, 3], [], None, None)
    assert tuple(out.shape) == (3, 0)
    out, st = dev.snp_cross_kernel(s, s, [1, 2, 3], [4, 5], [], [])
    assert tuple(out.shape) == (3, 2) and float(out.abs().sum()) == 0.0
    with pytest.raises(ValueError):
        dev.snp_cross_kernel(s, s, None, None, [1, 2], [1])


@pytest.mark.parametrize("pinned", [False, True])
def test_snp_kernel_host_entry(pinned, oracle, dev, monkeypatch):
    """pstb_snp_kernel_host: packed host bytes -> K on the host (float32 / float64), several H2D slices, gathered axes, trained stats."""
    import ctypes
    from pysnptools_b200._lib import lib, check, F32, F64, STD_UNIT, STD_BETA
    n, m = 700, 1500
    rec = (n + 3) // 4
    packed = oracle.synth_packed(n, 0, m, missing_rate=0.03, seed=11)
    monkeypatch.setenv("PSTB_KERNEL_SLICE_SNPS", "256")                       # 6 slices of 256 SNPs (4 chunks of 64 each)
    if pinned:
        p_in, p_out = lib.pstb_host_alloc(m * rec), lib.pstb_host_alloc(n * n * 8)
        h_packed = np.ctypeslib.as_array(ctypes.cast(p_in, ctypes.POINTER(ctypes.c_uint8)), shape=(m, rec))
        h_packed[...] = packed
    else:
        h_packed = packed
    rng = np.random.default_rng(5)
    ii = rng.permutation(n)[:300].astype(np.int64)
    si = rng.permutation(m)[:800].astype(np.int64)
    for dtype, code in ((np.float64, F64), (np.float32, F32)):
        for (mode, args, ab) in ((STD_UNIT, {}, (float("nan"), float("nan"))), (STD_BETA, dict(is_beta=True, a=1, b=25), (1.0, 25.0))):
            for iid_idx, sid_idx in ((None, None), (ii, si)):
                ni, ns = (n if iid_idx is None else len(iid_idx)), (m if sid_idx is None else len(sid_idx))
                ref, rst = oracle.read_kernel(packed, n, iid_index=iid_idx, sid_index=sid_idx, **args)
                if pinned:
                    K = np.ctypeslib.as_array(ctypes.cast(p_out, ctypes.POINTER(ctypes.c_double if dtype == np.float64 else ctypes.c_float)), shape=(ni, ni))
                else:
                    K = np.empty((ni, ni), dtype=dtype)
                K[...] = -1
                stats = np.empty((ns, 2))
                check(lib.pstb_snp_kernel_host(h_packed.ctypes.data, n, m, None if iid_idx is None else iid_idx.ctypes.data, ni,
                                               None if sid_idx is None else sid_idx.ctypes.data, ns, 0, mode, ab[0], ab[1], 0,
                                               stats.ctypes.data, K.ctypes.data, code, 64, -1))
                assert rel_fro(K.astype(np.float64), ref) < K_TOL and np.array_equal(K, K.T)
                np.testing.assert_allclose(stats, rst, rtol=1e-12)
                K2 = np.empty((ni, ni), dtype=dtype)
                check(lib.pstb_snp_kernel_host(h_packed.ctypes.data, n, m, None if iid_idx is None else iid_idx.ctypes.data, ni,
                                               None if sid_idx is None else sid_idx.ctypes.data, ns, 0, mode, ab[0], ab[1], 1,
                                               stats.ctypes.data, K2.ctypes.data, code, 128, -1))
                assert rel_fro(K2.astype(np.float64), ref) < K_TOL
    # no SNPs -> zeros; bad index -> error
    Kz = np.full((n, n), 3.0, dtype=np.float32)
    st0 = np.empty((0, 2))
    empty = np.zeros(0, dtype=np.int64)
    check(lib.pstb_snp_kernel_host(h_packed.ctypes.data, n, m, None, n, empty.ctypes.data, 0, 0, STD_UNIT, 0.0, 0.0, 0, st0.ctypes.data, Kz.ctypes.data, F32, 64, -1))
    assert not Kz.any()
    bad = np.array([m], dtype=np.int64)
    assert lib.pstb_snp_kernel_host(h_packed.ctypes.data, n, m, None, n, bad.ctypes.data, 1, 0, STD_UNIT, 0.0, 0.0, 0, np.empty((1, 2)).ctypes.data, Kz.ctypes.data, F32, 64, -1) != 0
    if pinned:
        del h_packed, K
        lib.pstb_host_free(p_in)
        lib.pstb_host_free(p_out)


@pytest.mark.parametrize("n,m", [(300, 200), (700, 300), (1500, 130)])
def test_kernel_from_tiles_matches_square_path(n, m, oracle, dev):
    """Compact lower-triangular tiles -> full K (pstb_kernel_from_tiles): what the multi-GPU path does after all-reducing the
    triangle.  Two SNP shards summed in tile storage == the one-pass square kernel; the result is exactly symmetric."""
    import torch
    packed = oracle.synth_packed(n, 0, m, missing_rate=0.03, seed=n)
    store = dev.PackedStore.from_host(packed, n)
    ref, _ = oracle.read_kernel(packed, n)
    half = m // 2
    # each "rank" standardizes its own SNPs (statistics are per SNP, so shards agree with the one-pass kernel)
    t0, _c, _s = dev.snp_kernel_tiles(store, None, slice(0, half), chunk=64)
    t1, _c, _s = dev.snp_kernel_tiles(store, None, slice(half, m), chunk=64)
    K = dev.kernel_from_tiles(t0 + t1, n)
    Kc = K.double().cpu().numpy()
    assert np.array_equal(Kc, Kc.T) and rel_fro(Kc, ref) < K_TOL
    Ks, _ = dev.snp_kernel(store, chunk=64)
    assert rel_fro(Kc, Ks.double().cpu().numpy()) < 2e-6
    # tiles of two ranks expanded into one matrix cover everything once
    Kr = torch.full((n, n), float("nan"), device="cuda")
    for rank in range(2):
        tr, _c, _s = dev.snp_kernel_tiles(store, rank=rank, world=2, chunk=64)
        dev.kernel_from_tiles(tr, n, rank=rank, world=2, K=Kr)
    Krc = Kr.double().cpu().numpy()
    assert not np.isnan(Krc).any() and rel_fro(Krc, ref) < K_TOL


@pytest.mark.parametrize("n,m", [(515, 300), (700, 4096), (2100, 700)])
@pytest.mark.parametrize("missing", [0.0, 0.05])
def test_low_term_modes(n, m, missing, oracle, dev):
    """Exact-dosage path, with and without missing genotypes: the low term on the fp8 pipe (e4m3 x e4m3) stays inside the 1e-5 gate
    for M >> N, M ~ N and M << N; the fp16 low term lands near 1e-6; 'auto' picks fp8 when a call has at least as many SNPs as
    individuals (4 x as many for Beta).  The mode is a per-call argument; the process-wide default only feeds 'default'."""
    packed = oracle.synth_packed(n, 0, m, missing_rate=missing, seed=n + m)
    store = dev.PackedStore.from_host(packed, n)
    assert dev.get_syrk_low_term() == "auto"
    errs = {}
    for std, args in ((("unit",), {}), (("beta", 1, 25), dict(is_beta=True, a=1, b=25))):
        ref, _ = oracle.read_kernel(packed, n, **args)
        for mode in ("fp16", "fp8", "auto", "default"):
            K, _ = dev.snp_kernel(store, standardizer=std, chunk=256, low_term=mode)
            Kc = K.double().cpu().numpy()
            assert np.array_equal(Kc, Kc.T)
            errs[(std[0], mode)] = rel_fro(Kc, ref)
    assert all(e < K_TOL for e in errs.values()), errs
    assert all(errs[(s, "fp16")] < 3e-6 for s in ("unit", "beta")), errs
    for s, need in (("unit", n), ("beta", 4 * n)):
        want_auto = "fp8" if m >= need else "fp16"
        assert errs[(s, "auto")] == errs[(s, want_auto)] == errs[(s, "default")], errs
        assert dev.low_term_for(m, n, (s,)) == want_auto
    assert dev.low_term_for(10 * n, n) == "fp8"                             # a sharded kernel with many SNPs in total
    try:
        assert dev.set_syrk_low_term("fp16") == "auto"                      # the process default feeds 'default' only
        K, _ = dev.snp_kernel(store, chunk=256)
        assert rel_fro(K.double().cpu().numpy(), oracle.read_kernel(packed, n)[0]) == errs[("unit", "fp16")]
        K, _ = dev.snp_kernel(store, chunk=256, low_term="fp8")
        assert rel_fro(K.double().cpu().numpy(), oracle.read_kernel(packed, n)[0]) == errs[("unit", "fp8")]
    finally:
        dev.set_syrk_low_term("auto")


def test_exact_dosage_path_with_missing_and_trained_stats(oracle, dev, monkeypatch):
    """Round 2: chunks with missing genotypes and trained statistics take the 2-term exact-dosage GEMM (missing entries of the left
    plane hold fp16(mu - mu'), so mean imputation stays exact).  Checks: the heavy-missing case against the float64 oracle, that
    the result differs from the forced 3-term split only by rounding, trained statistics of another iid set, an all-missing SNP,
    an SNC SNP, and a trained mean outside [0, 2] (falls back to the 3-term split on the device)."""
    n, m = 900, 1400
    packed = oracle.synth_packed(n, 0, m, missing_rate=0.3, seed=11)
    packed[5, :] = 0x55                                                     # SNP 5: every genotype missing
    packed[9, :] = 0x00                                                     # SNP 9: constant (SNC)
    store = dev.PackedStore.from_host(packed, n)
    for std, args in ((("unit",), {}), (("beta", 1, 25), dict(is_beta=True, a=1, b=25))):
        ref, rst = oracle.read_kernel(packed, n, **args)
        for lt in ("fp16", "fp8"):
            K, st = dev.snp_kernel(store, standardizer=std, chunk=512, low_term=lt)
            Kc = K.double().cpu().numpy()
            assert not np.isnan(Kc).any() and np.array_equal(Kc, Kc.T)
            assert rel_fro(Kc, ref) < (3e-6 if lt == "fp16" else K_TOL), (std, lt, rel_fro(Kc, ref))
        np.testing.assert_allclose(st.cpu().numpy(), rst, rtol=1e-12, equal_nan=True)
    # trained statistics from a different iid subset, applied to the other individuals (UnitTrained, unittrained.py:47-70)
    train, test = np.arange(0, 600), np.arange(600, n)
    _, st_train = oracle.read_kernel(packed, n, iid_index=train)
    x_t, _ = oracle.standardize(oracle.decode(packed, n, test), use_stats=True, stats=st_train)
    ref_t = x_t @ x_t.T
    Kt, _ = dev.snp_kernel(store, test, None, stats=st_train, chunk=512, low_term="fp16")
    assert rel_fro(Kt.double().cpu().numpy(), ref_t) < 3e-6
    # the forced 3-term split agrees
    monkeypatch.setenv("PSTB_SYRK_3TERM", "1")
    K3, _ = dev.snp_kernel(store, chunk=512)
    monkeypatch.delenv("PSTB_SYRK_3TERM")
    K2, _ = dev.snp_kernel(store, chunk=512, low_term="fp16")
    ref, _ = oracle.read_kernel(packed, n)
    assert rel_fro(K3.double().cpu().numpy(), ref) < 3e-6 and rel_fro(K2.double().cpu().numpy(), K3.double().cpu().numpy()) < 3e-6
    # a trained mean outside [0, 2] cannot be centred exactly in fp16: that chunk takes the 3-term split, the others do not
    st_odd = oracle.read_kernel(packed, n)[1].copy()
    st_odd[700] = (3.75, 0.5)
    st_odd[20] = (-0.3, 2.0)
    x_o, _ = oracle.standardize(oracle.decode(packed, n), use_stats=True, stats=st_odd)
    ref_o = x_o @ x_o.T
    Ko, _ = dev.snp_kernel(store, stats=st_odd, chunk=512, low_term="fp8")
    assert rel_fro(Ko.double().cpu().numpy(), ref_o) < K_TOL


def test_low_term_is_per_call_under_threads(oracle, dev):
    """Two Python threads run kernels with different per-call modes at the same time (the C ABI promises concurrent callers on
    different streams): every result equals the single-threaded result of its own mode bit for bit."""
    import threading
    import torch
    n, m = 600, 2048
    packed = oracle.synth_packed(n, 0, m, missing_rate=0.02, seed=3)
    store = dev.PackedStore.from_host(packed, n)
    want = {lt: dev.snp_kernel(store, chunk=256, low_term=lt)[0].cpu().numpy() for lt in ("fp16", "fp8")}
    assert not np.array_equal(want["fp16"], want["fp8"])
    torch.cuda.synchronize()
    bad = []

    def worker(lt):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(6):
                K, _ = dev.snp_kernel(store, chunk=256, low_term=lt)
                s.synchronize()
                if not np.array_equal(K.cpu().numpy(), want[lt]):
                    bad.append(lt)

    th = [threading.Thread(target=worker, args=(lt,)) for lt in ("fp16", "fp8")]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not bad, bad


def test_float64_kernel_path(oracle, dev):
    """K3 in float64 (syrk_f64.cu): the dtype=float64 contract of the reference (val.dot(val.T) is a DGEMM then, snpdata.py:203-206; its
    tests compare float64 kernels to 10 decimals).  Fused decode + standardize into a float64 panel + fp64-FMA SYRK: ~1e-14 relative to the
    float64 oracle, for missing data, Beta, gathered axes, trained statistics, ragged chunks, accumulation over SNP halves; plus the
    float-matrix and host-buffer entry points."""
    import ctypes
    import torch
    from pysnptools_b200 import _lib
    n, m = 700, 1500
    packed = oracle.synth_packed(n, 0, m, missing_rate=0.05, seed=21)
    packed[7, :] = 0x55                                                   # an all-missing SNP and an SNC one
    packed[8, :] = 0xFF
    store = dev.PackedStore.from_host(packed, n)
    for std, args in ((("unit",), {}), (("beta", 1, 25), dict(is_beta=True, a=1, b=25))):
        ref, rst = oracle.read_kernel(packed, n, **args)
        for chunk in (None, 100, 64):
            K, st = dev.snp_kernel_f64(store, standardizer=std, chunk=chunk)
            Kc = K.cpu().numpy()
            assert K.dtype == torch.float64 and np.array_equal(Kc, Kc.T) and not np.isnan(Kc).any()
            assert rel_fro(Kc, ref) < 1e-13, (std, chunk, rel_fro(Kc, ref))
            assert np.max(np.abs(Kc - ref)) < 1e-10 * max(1.0, np.max(np.abs(ref)) / 1e3)      # the reference's "10 decimals" at these magnitudes
        np.testing.assert_allclose(st.cpu().numpy(), rst, rtol=1e-12, equal_nan=True)
    rng = np.random.default_rng(2)
    ii, si = rng.permutation(n)[:333], rng.permutation(m)[:1000]
    ref, rst = oracle.read_kernel(packed, n, iid_index=ii, sid_index=si, count_A1=True)
    K, st = dev.snp_kernel_f64(store, ii, si, count_A1=True, chunk=256)
    assert rel_fro(K.cpu().numpy(), ref) < 1e-13
    K2, _ = dev.snp_kernel_f64(store, ii, si, count_A1=True, stats=st, chunk=256)              # trained statistics
    assert rel_fro(K2.cpu().numpy(), ref) < 1e-13
    Ka, _ = dev.snp_kernel_f64(store, ii, si[:500], count_A1=True, chunk=128, mirror=False)
    Ka, _ = dev.snp_kernel_f64(store, ii, si[500:], count_A1=True, chunk=128, K=Ka, accumulate=True)
    assert rel_fro(Ka.cpu().numpy(), ref) < 1e-13
    # float matrices, C and F order
    v = rng.standard_normal((130, 777))
    want = v.dot(v.T)
    for arr in (np.ascontiguousarray(v), np.asfortranarray(v)):
        t = torch.from_numpy(arr).cuda() if arr.flags["C_CONTIGUOUS"] else torch.from_numpy(arr.T).cuda().t()
        Kf = dev.float_kernel_f64(t).cpu().numpy()
        assert rel_fro(Kf, want) < 1e-14 and np.array_equal(Kf, Kf.T)
    # host-buffer entry point, several slices
    lib, check = _lib.lib, _lib.check
    os.environ["PSTB_KERNEL_SLICE_SNPS"] = "512"
    try:
        ref, rst = oracle.read_kernel(packed, n)
        Kh = np.empty((n, n))
        sth = np.empty((m, 2))
        check(lib.pstb_snp_kernel_host_f64(packed.ctypes.data, n, m, None, n, None, m, 0, _lib.STD_UNIT, float("nan"), float("nan"), 0,
                                           sth.ctypes.data, Kh.ctypes.data, 128))
        assert rel_fro(Kh, ref) < 1e-13 and np.array_equal(Kh, Kh.T)
        np.testing.assert_allclose(sth, rst, rtol=1e-12, equal_nan=True)
    finally:
        del os.environ["PSTB_KERNEL_SLICE_SNPS"]
    _lib.lib.pstb_host_release()


@pytest.mark.parametrize("n,m,chunk,bands", [(700, 1000, 256, 4), (1300, 200, 256, 3), (515, 512, 256, 8), (300, 0, 64, 2)])
def test_banded_last_chunk_matches_plain_kernel(n, m, chunk, bands, oracle, dev):
    """The overlapped multi-GPU reduction multiplies the last SNP chunks band-major (pstb_snp_kernel_tiles_band: planes of 1-3 tail chunks
    built once, every band of tiles takes them in turn) and expands finished bands on a side stream (pstb_kernel_from_tiles_range).  On one
    GPU (no collective) the result must equal the plain kernel: same planes, same tiles, only the rank-one part is added per call."""
    import torch
    from pysnptools_b200 import parallel
    packed = oracle.synth_packed(n, 0, m, missing_rate=0.03, seed=n + m) if m else np.zeros((0, (n + 3) // 4), dtype=np.uint8)
    store = dev.PackedStore.from_host(packed, n)
    for std, tail in ((("unit",), 1), (("unit",), 3), (("beta", 1, 25), 2)):
        K, st = parallel.snp_kernel_sharded_overlapped(store, n, m, None, std, chunk=chunk, bands=bands, tail_chunks=tail)
        torch.cuda.synchronize()
        Kp, stp = dev.snp_kernel(store, standardizer=std, chunk=chunk, low_term=dev.low_term_for(m, n, std))
        Kc, Kpc = K.double().cpu().numpy(), Kp.double().cpu().numpy()
        assert np.array_equal(Kc, Kc.T)
        if m:
            assert rel_fro(Kc, Kpc) < 3e-7, rel_fro(Kc, Kpc)
            assert np.array_equal(st.cpu().numpy(), stp.cpu().numpy())
            # the default multi-GPU path on one GPU: compact tiles with the rank-one vector DEFERRED (all-reduced as n doubles on N GPUs)
            # and added by the expansion, in slices
            tiles, _c, st2, u = dev.snp_kernel_tiles(store, standardizer=std, chunk=chunk, low_term=dev.low_term_for(m, n, std), defer_rank1=True)
            Kd = torch.full((n, n), float("nan"), dtype=torch.float32, device="cuda")
            parallel.allreduce_tiles_and_expand(tiles, n, Kd, slices=3, u=u)
            assert np.array_equal(Kd.cpu().numpy(), Kp.cpu().numpy())          # same sums, same single rounding of tile + v_k
            assert u.dtype == torch.float64 and tuple(u.shape) == (n,)
            ref, _ = oracle.read_kernel(packed, n, **({} if std[0] == "unit" else dict(is_beta=True, a=1, b=25)))
            assert rel_fro(Kc, ref) < K_TOL
        else:
            assert not Kc.any()


@pytest.mark.parametrize("pinned", [False, True])
@pytest.mark.parametrize("n,m,slice_snps", [(1000, 3000, 512), (1300, 900, 256), (2600, 2200, 1024)])
def test_snp_kernel_host_overlapped_copy_out(pinned, n, m, slice_snps, oracle, dev, monkeypatch):
    """pstb_snp_kernel_host with the copy-out of K overlapped (compact tiles, band-major tail from the bottom band up, finished row
    ranges leave while the bands above multiply; PSTB_HOST_KERNEL_OVERLAP=2 forces the path at test sizes) against the plain path
    (=0) and the oracle: head of several slices + tail, tail only (m = 900: every chunk is a tail chunk), gathered individuals and
    scattered SNPs, float32 / float64 output, Unit / Beta with missing genotypes, trained statistics."""
    import ctypes
    from pysnptools_b200._lib import lib, check, F32, F64, STD_UNIT, STD_BETA
    rec = (n + 3) // 4
    packed = oracle.synth_packed(n, 0, m, missing_rate=0.03, seed=n + 1)
    monkeypatch.setenv("PSTB_KERNEL_SLICE_SNPS", str(slice_snps))
    held = []
    if pinned:
        p_in = lib.pstb_host_alloc(m * rec)
        held.append(p_in)
        h_packed = np.ctypeslib.as_array(ctypes.cast(p_in, ctypes.POINTER(ctypes.c_uint8)), shape=(m, rec))
        h_packed[...] = packed
    else:
        h_packed = packed
    rng = np.random.default_rng(n)
    ii = np.sort(rng.permutation(n)[: n - 130]).astype(np.int64)[::-1].copy()
    si = rng.permutation(m)[: m - 77].astype(np.int64)

    def call(mode, ab, iid_idx, sid_idx, code, dtype, use_stats=0, stats=None, chunk=64):
        ni, ns = (n if iid_idx is None else len(iid_idx)), (m if sid_idx is None else len(sid_idx))
        if pinned:
            p_out = lib.pstb_host_alloc(ni * ni * 8)
            held.append(p_out)
            K = np.ctypeslib.as_array(ctypes.cast(p_out, ctypes.POINTER(ctypes.c_double if dtype == np.float64 else ctypes.c_float)), shape=(ni, ni))
        else:
            K = np.empty((ni, ni), dtype=dtype)
        K[...] = -7
        st = np.empty((ns, 2)) if stats is None else stats
        check(lib.pstb_snp_kernel_host(h_packed.ctypes.data, n, m, None if iid_idx is None else iid_idx.ctypes.data, ni,
                                       None if sid_idx is None else sid_idx.ctypes.data, ns, 0, mode, ab[0], ab[1], use_stats,
                                       st.ctypes.data, K.ctypes.data, code, chunk, -1))
        return K, st

    try:
        for dtype, code in ((np.float32, F32), (np.float64, F64)):
            for (mode, args, ab) in ((STD_UNIT, {}, (float("nan"), float("nan"))), (STD_BETA, dict(is_beta=True, a=1, b=25), (1.0, 25.0))):
                for iid_idx, sid_idx in ((None, None), (ii, si)):
                    ref, rst = oracle.read_kernel(packed, n, iid_index=iid_idx, sid_index=sid_idx, **args)
                    monkeypatch.setenv("PSTB_HOST_KERNEL_OVERLAP", "0")
                    K0, st0 = call(mode, ab, iid_idx, sid_idx, code, dtype)
                    K0 = K0.copy()
                    monkeypatch.setenv("PSTB_HOST_KERNEL_OVERLAP", "2")
                    K1, st1 = call(mode, ab, iid_idx, sid_idx, code, dtype)
                    assert np.array_equal(K1, K1.T) and np.isfinite(K1).all()
                    assert rel_fro(K1.astype(np.float64), ref) < K_TOL
                    np.testing.assert_allclose(st1, rst, rtol=1e-12)
                    assert np.array_equal(st0, st1)
                    # the same tiles, the same chunk order; only the rank-one vectors are summed per call instead of in one buffer
                    assert rel_fro(K1.astype(np.float64), K0.astype(np.float64)) < 2e-7
                    K2, _ = call(mode, ab, iid_idx, sid_idx, code, dtype, use_stats=1, stats=st1.copy(), chunk=128)
                    assert rel_fro(K2.astype(np.float64), ref) < K_TOL
    finally:
        for ptr in held:
            lib.pstb_host_free(ptr)
