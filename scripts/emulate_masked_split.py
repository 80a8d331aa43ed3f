"""CPU emulation of the K3 operand split with MISSING genotypes ("exact-dosage path"), run before the
kernel was changed.  Not product code; needs no GPU.

    true left   T_ij = m_ij (g_ij - mu_j)            (0 for a missing genotype)
    left plane  L_ij = g_ij - mu'_j   (observed; mu' = mu rounded to 10 fractional bits: exact in fp16)
                     = fp16(-delta_j) (missing;  delta = mu' - mu)            => T = L + delta up to 2^-22
    right plane B_kj = w_j m_kj (g_kj - mu_j) = hi (fp16) + lo (fp16 | e4m3)
    K_ik = sum_j L_ij B_kj + v_k,   v_k = sum_j delta_j B_kj  (fp64, rank one)

Prints the relative Frobenius error of that scheme against float64 X X^T for Unit / Beta, with the low term in fp16 and in
e4m3 (tensor products are exact in fp32; accumulation is emulated in float64, i.e. this isolates operand rounding).
"""
import sys

import numpy as np


def e4m3(x):
    """round-to-nearest-even to e4m3 (4 exponent bits, bias 7, 3 mantissa bits; max 448; subnormals at 2^-9), saturating."""
    x = np.asarray(x, dtype=np.float64)
    s = np.sign(x)
    a = np.minimum(np.abs(x), 448.0)
    e = np.floor(np.log2(np.where(a > 0, a, 1.0)))
    e = np.maximum(e, -6.0)                      # subnormal range shares the exponent of the smallest normal
    q = np.exp2(e - 3)
    return s * np.rint(a / q) * q


def synth(n, m, missing, seed):
    rng = np.random.default_rng(seed)
    p = rng.uniform(0.05, 0.5, size=(1, m))
    g = (rng.random((n, m)) < p).astype(np.float64) + (rng.random((n, m)) < p)
    g[rng.random((n, m)) < missing] = np.nan
    return g


def beta_pdf(maf, a, b):
    from math import lgamma
    lnB = lgamma(a) + lgamma(b) - lgamma(a + b)
    return np.exp((a - 1) * np.log(maf) + (b - 1) * np.log1p(-maf) - lnB)


def run(n, m, missing, mode, low, seed=0, count_major=False):
    g = synth(n, m, missing, seed)
    if count_major:
        g = 2.0 - g
    obs = ~np.isnan(g)
    mu = np.nanmean(g, axis=0)
    sd = np.nanstd(g, axis=0)
    ok = sd > 0
    if mode == "unit":
        f = np.where(ok, 1.0 / np.where(ok, sd, 1.0), 0.0)
    else:
        maf = mu / 2
        maf = np.where(maf > 0.5, 1 - maf, maf)
        f = np.where(ok, beta_pdf(np.clip(maf, 1e-300, 1), 1.0, 25.0), 0.0)
    T = np.where(obs, g - mu, 0.0)
    X = T * f
    K = X @ X.T
    w = f * f
    mu_r = np.rint(mu * 1024) / 1024
    delta = mu_r - mu
    L = np.where(obs, np.nan_to_num(g) - mu_r, np.float16(-delta).astype(np.float64))
    assert np.array_equal(L, L.astype(np.float16).astype(np.float64)), "left plane is not exact in fp16"
    B = w * T
    amax = np.max(np.abs(B))
    scale = 2.0 ** (14 - np.ceil(np.log2(amax)))
    Bs = B * scale
    hi = Bs.astype(np.float16).astype(np.float64)
    res = Bs - hi
    if low == "fp16":
        lo = res.astype(np.float16).astype(np.float64)
        Ll = L
    elif low == "fp8":
        lo = e4m3(res)
        Ll = e4m3(L)
    else:
        lo = 0 * res
        Ll = L
    v = (delta * B).sum(axis=1)
    Kh = (L @ hi.T + Ll @ lo.T) / scale + v[None, :]
    low_tri = np.tril(np.ones((n, n), dtype=bool))
    Kh = np.where(low_tri, Kh, Kh.T)
    return np.linalg.norm(Kh - K) / np.linalg.norm(K)


if __name__ == "__main__":
    shapes = [(512, 5120), (1024, 2048), (1024, 1024), (2048, 512)]
    if len(sys.argv) > 1:
        shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
    print("%-12s %-5s %-7s %-6s %10s %10s %10s" % ("shape", "mode", "missing", "major", "hi only", "fp16 lo", "e4m3 lo"))
    for n, m in shapes:
        for mode in ("unit", "beta"):
            for missing in (0.0, 0.05, 0.3):
                for major in (False, True):
                    errs = [run(n, m, missing, mode, low, count_major=major) for low in ("none", "fp16", "fp8")]
                    print("%-12s %-5s %-7.2f %-6s %10.2e %10.2e %10.2e" % ("%dx%d" % (n, m), mode, missing, major, *errs), flush=True)
