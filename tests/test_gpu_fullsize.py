"""GPU: BASELINE.json's full sizes, checked through size-independent properties + sampled oracle comparisons.

cfg2  10 000 x 1 000 000 decode + Unit f32  : per-SNP sums ~ 0, sums of squares = n_obs, sampled columns == oracle
cfg4  100 000 x 200 000, 5 % missing, Beta(1,25), random 1/2 x 1/2 gather: sampled columns == oracle (same gather)
cfg3  50 000 x (a 16 384-SNP slice of 500 000) SnpKernel(Unit): trace(K) = N*M, symmetry, sampled entries == oracle
"""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def env():
    import torch
    import bench
    from pysnptools_b200 import device
    return torch, bench, device


def test_cfg2_full_size_properties(env, oracle):
    torch, bench, dev = env
    n, m = 10_000, 1_000_000
    store = bench.gen_store_device(dev, torch, n, m, seed=11)
    val, stats = dev.read(store, dtype=np.float32, order="F", standardizer=("unit",))
    assert tuple(val.shape) == (n, m) and val.t().is_contiguous()
    base = val.t()                                                     # [m, n] contiguous
    col_sum = base.sum(dim=1, dtype=torch.float64)
    col_sq = torch.cat([(base[s:s + 50000].double() ** 2).sum(dim=1) for s in range(0, m, 50000)])     # in slices: 40 GB of float64 temporaries otherwise
    sd = stats[:, 1]
    live = torch.isfinite(sd)
    assert float(col_sum.abs().max()) < 0.05                           # centred (f32 rounding of 10 000 terms)
    assert float((col_sq[live] / n - 1.0).abs().max()) < 1e-5          # unit variance, no missing data
    assert float(col_sq[~live].abs().max() if (~live).any() else 0.0) == 0.0
    assert not torch.isnan(val[:, ::1000]).any()
    rng = np.random.default_rng(0)
    cols = np.sort(rng.choice(m, size=64, replace=False))
    packed = store.tensor[torch.as_tensor(cols, device="cuda")][:, : (n + 3) // 4].cpu().numpy()
    ref, rst = oracle.standardize(oracle.decode(packed, n))
    got = val[:, torch.as_tensor(cols, device="cuda")].cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(stats[torch.as_tensor(cols, device="cuda")].cpu().numpy(), rst, rtol=1e-12)
    # decode of the same store is bit-exact on the sample, and idempotent across calls (checksum of checksums)
    raw, _ = dev.read(store, None, cols, dtype=np.int8, order="F")
    assert np.array_equal(raw.cpu().numpy(), oracle.decode(packed, n, dtype=np.int8))
    val2, _ = dev.read(store, dtype=np.float32, order="F", standardizer=("unit",))
    assert torch.equal(val2.t()[::997].view(torch.int32).sum(dim=1), base[::997].view(torch.int32).sum(dim=1))


def test_cfg4_full_size_gather_vs_oracle(env, oracle):
    torch, bench, dev = env
    n, m = 100_000, 200_000
    store = bench.gen_store_device(dev, torch, n, m, seed=12, missing_rate=0.05)
    rng = np.random.default_rng(1)
    ii = rng.permutation(n)[: n // 2]
    si = rng.permutation(m)[: m // 2]
    val, stats = dev.read(store, ii, si, dtype=np.float32, order="F", standardizer=("beta", 1, 25))
    assert tuple(val.shape) == (n // 2, m // 2) and not torch.isnan(val[:, ::500]).any()
    pick = np.sort(rng.choice(m // 2, size=48, replace=False))
    packed = store.tensor[torch.as_tensor(si[pick], device="cuda")][:, : (n + 3) // 4].cpu().numpy()
    ref, rst = oracle.standardize(oracle.decode(packed, n, ii), True, 1, 25)
    got = val[:, torch.as_tensor(pick, device="cuda")].cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(stats[torch.as_tensor(pick, device="cuda")].cpu().numpy(), rst, rtol=1e-12)
    frac_zero = float((val[:, torch.as_tensor(pick, device="cuda")] == 0).float().mean())
    assert 0.04 < frac_zero < 0.06                                     # the 5 % missing entries became 0


def test_cfg3_width_kernel_properties(env, oracle):
    torch, bench, dev = env
    n, m = 50_000, 16_384                                              # full cfg3 height, one operand-plane chunk of its SNPs
    store = bench.gen_store_device(dev, torch, n, m, seed=13)
    K, stats = dev.snp_kernel(store, chunk=8192)                       # two chunks: exercises the accumulate path at full height
    live = int(torch.isfinite(stats[:, 1]).sum())
    tr = float(K.diagonal().double().sum())
    assert abs(tr / (float(n) * live) - 1.0) < 5e-6                    # trace(K) = N * (#non-SNC SNPs) for Unit
    rows = torch.as_tensor(np.random.default_rng(2).choice(n, 64, replace=False), device="cuda")
    assert torch.equal(K[rows][:, rows], K[rows][:, rows].t())         # symmetric after the mirror
    ii = np.sort(rows.cpu().numpy())
    packed = store.tensor[:, : (n + 3) // 4].cpu().numpy()
    x, _ = oracle.standardize(oracle.decode(packed, n)[ii])            # NB: statistics must come from all iids
    _, st_all = oracle.standardize(oracle.decode(packed[:256], n))
    np.testing.assert_allclose(stats[:256].cpu().numpy(), st_all, rtol=1e-12)
    xs, _ = oracle.standardize(oracle.decode(packed, n, ii), use_stats=True, stats=stats.cpu().numpy())
    ref = xs @ xs.T
    got = K[torch.as_tensor(ii, device="cuda")][:, torch.as_tensor(ii, device="cuda")].double().cpu().numpy()
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-5


@pytest.mark.parametrize("spec,m,missing,want_low", [(("unit",), 50_048, 0.0, "fp8"), (("unit",), 50_048, 0.05, "fp8"), (("beta", 1, 25), 50_048, 0.05, "fp16"),
                                                       (("beta", 1, 25), 200_192, 0.0, "fp8")])
def test_cfg3_height_kernel_against_oracle_on_sampled_tiles(env, spec, m, missing, want_low):
    """The K path the benchmark times, at cfg3's FULL height (N = 50 000) and with at least as many SNPs as individuals, so that AUTO takes
    the fp8 low term exactly as it does for cfg3 (round 1 only tested it up to N = 2 100): 15 sampled 256 x 256 tiles (all pairs of 5 row
    blocks, the ragged last one among them) against the float64 oracle over ALL SNPs -- per-tile and stratified whole-matrix relative
    Frobenius error under the 1e-5 gate, diagonal bias under 3e-6 -- for Unit with and without missing genotypes and for Beta(1,25)
    (fp16 low term below 4 N SNPs, fp8 above)."""
    torch, bench, dev = env
    n = 50_000
    assert dev.low_term_for(m, n, spec) == want_low
    store = bench.gen_store_device(dev, torch, n, m, seed=31 + m, missing_rate=missing)
    K, stats = dev.snp_kernel(store, standardizer=spec)
    blocks = bench.pick_blocks(n, 5, seed=7)

    def fetch(I, J):
        sub = K[I * 256:I * 256 + 256, J * 256:J * 256 + 256].double().cpu().numpy()
        out = np.zeros((256, 256))
        out[: sub.shape[0], : sub.shape[1]] = sub
        return out
    res = bench.sampled_tile_parity(torch, None, 1, 0, bench._oracle_lib(), store, stats, n, spec, fetch, blocks)
    assert res["stats_match_oracle_rtol_1e-12"] and res["sampled_tiles"] == 15
    assert res["worst_rel_frobenius_vs_oracle"] < 1e-5, res
    assert res["rel_frobenius_whole_K_estimate"] < 6e-6, res
    assert abs(res["diag_rel_bias"]) < 3e-6, res
    assert torch.equal(K[:300, -300:], K[-300:, :300].t())
