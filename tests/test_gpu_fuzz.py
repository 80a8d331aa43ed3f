"""GPU: seeded randomized sweep over shapes / selections / dtypes / orders / standardizers against the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sel(rng, count):
    kind = rng.integers(0, 7)
    if kind == 0 or count == 0:
        return None, np.arange(count)
    if kind == 1:
        s = slice(None, None, -1)
        return s, np.arange(count)[s]
    if kind == 2:
        a, b, st = int(rng.integers(0, count)), int(rng.integers(0, count + 1)), int(rng.integers(1, 5))
        s = slice(a, b, st)
        return s, np.arange(count)[s]
    if kind == 3:
        idx = rng.permutation(count)[: int(rng.integers(1, count + 1))]
        return idx, idx
    if kind == 4:
        idx = rng.integers(0, count, size=int(rng.integers(1, count + 5)))       # repeats
        return idx, idx
    if kind == 5:
        idx = np.sort(rng.permutation(count)[: int(rng.integers(1, count + 1))])
        return idx, idx
    start = int(rng.integers(0, count)) // 16 * 16
    s = slice(start, None)
    return s, np.arange(count)[s]


@pytest.mark.parametrize("seed", range(6))
def test_randomized_reads_and_kernels(seed, oracle):
    import torch
    from pysnptools_b200 import device as dev
    rng = np.random.default_rng(1000 + seed)
    for case in range(18):
        n = int(rng.choice([1, 2, 3, 5, 17, 64, 100, 255, 256, 257, 1000, 2049, 5003]))
        m = int(rng.choice([1, 2, 7, 31, 64, 65, 130]))
        miss = float(rng.choice([0.0, 0.0, 0.05, 0.5]))
        packed = oracle.synth_packed(n, 0, m, miss, seed=seed * 100 + case)
        if rng.random() < 0.3:
            store = dev.PackedStore(torch.from_numpy(packed).cuda(), n, m)       # tight ld: unaligned paths
        else:
            store = dev.PackedStore.from_host(packed, n)
        isel, ii = _sel(rng, n)
        ssel, si = _sel(rng, m)
        a1 = bool(rng.integers(0, 2))
        dtype = [np.float32, np.float64, np.int8][int(rng.integers(0, 3))]
        order = "FC"[int(rng.integers(0, 2))]
        ctx = (seed, case, n, m, miss, a1, dtype.__name__, order)
        val, _ = dev.read(store, isel, ssel, count_A1=a1, dtype=dtype, order=order)
        ref = oracle.decode(packed, n, ii, si, a1, dtype, order)
        assert val.shape == ref.shape and np.array_equal(val.cpu().numpy(), ref, equal_nan=dtype != np.int8), ctx
        if len(ii) == 0 or len(si) == 0:
            continue
        fdt = np.float32 if dtype == np.int8 else dtype
        std, args = [(("unit",), (False, np.nan, np.nan)), (("beta", 1, 25), (True, 1, 25)), (("beta", 0.5, 3.0), (True, 0.5, 3.0))][int(rng.integers(0, 3))]
        raw = oracle.decode(packed, n, ii, si, a1)
        want, wst = oracle.standardize(raw, *args)
        got, st = dev.read(store, isel, ssel, count_A1=a1, dtype=fdt, order=order, standardizer=std)
        ok_cols = np.isfinite(want).all(axis=0)                                    # Beta(a<1) at maf 0 is inf * 0 in the python twin: unpinned
        np.testing.assert_allclose(got.cpu().numpy()[:, ok_cols], want[:, ok_cols], rtol=1e-6, atol=1e-6 if fdt == np.float32 else 1e-11, err_msg=str(ctx))
        np.testing.assert_allclose(st.cpu().numpy(), wst, rtol=1e-12, err_msg=str(ctx))
        if std[0] == "beta" and std[1] < 1:
            continue
        K, _ = dev.snp_kernel(store, isel, ssel, count_A1=a1, standardizer=std, chunk=64)
        x = np.where(np.isfinite(want), want, 0.0)
        Kref = x @ x.T
        nrm = np.linalg.norm(Kref)
        if nrm > 0:
            assert np.linalg.norm(K.double().cpu().numpy() - Kref) / nrm < 1e-5, ctx
