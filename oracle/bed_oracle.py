"""CPU oracle for the genotype hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``pysnptools_b200``) never routes through it and fails loudly without the CUDA
library.

It is a plain-NumPy restatement of what the reference computes on the path
decode(.bed) -> standardize(Unit|Beta) -> K = X X^T:

* decode:       the 2-bit PLINK layout the reference reads through
                ``bed_reader.open_bed(...).read`` (reference call site
                ``pysnptools/snpreader/bed.py:318-345``; dtypes / int8 missing value
                ``bed.py:54-58``; count_A1 meaning ``bed.py:27``).  ``bed_reader`` is a
                third-party Rust wheel (``bed-reader>=0.2.36``, ``setup.py:27``) that is
                not vendored in the reference tree; its decode is the published PLINK
                format.
* standardize:  ``pysnptools/standardizer/standardizer.py:135-163`` (Unit) and
                ``:175-211`` (Beta), the reference's own pure-Python twins of the Rust
                ``standardize_f32/f64``.
* kernel:       ``pysnptools/snpreader/snpreader.py:623-668`` (block loop) and
                ``pysnptools/snpreader/snpdata.py:190-214`` (``val.dot(val.T)``).

Parity pinning: ``tests/test_oracle_golden.py`` checks every function here against
the reference's own golden vectors (``tests/golden/``; produced by
``tests/golden/make_golden.py`` which imports the reference's Python layer).
"""
import math
import numpy as np

BED_MAGIC = bytes([0x6C, 0x1B, 0x01])

# code (2 bits, LSB pair first) -> value; code 1 is "missing"   
_LUT_A2 = (0.0, np.nan, 1.0, 2.0)   # count_A1=False
_LUT_A1 = (2.0, np.nan, 1.0, 0.0)   # count_A1=True
INT8_MISSING = -127                 # bed.py:54-56


def bytes_per_snp(iid_count):
    return (int(iid_count) + 3) // 4


def read_packed(path, iid_count, sid_count, skip_format_check=False):
    """Return the packed genotype records as uint8 [sid_count, ceil(iid_count/4)]."""
    b = bytes_per_snp(iid_count)
    with open(path, "rb") as f:
        head = f.read(3)
        if not skip_format_check and head != BED_MAGIC:
            raise ValueError("'{0}' is not a SNP-major .bed file (bad magic bytes)".format(path))
        raw = np.fromfile(f, dtype=np.uint8)
    if raw.size != sid_count * b:
        raise ValueError("'{0}': expected {1} genotype bytes, found {2}".format(path, sid_count * b, raw.size))
    return raw.reshape(sid_count, b)


def count_lines(path):
    n = 0
    with open(path, "rb") as f:
        for _ in f:
            n += 1
    return n


def unpack_codes(packed, iid_count):
    """uint8 codes [sid, iid] in {0,1,2,3}; pair k of byte q is individual 4q+k."""
    m, b = packed.shape
    codes = np.empty((m, b, 4), dtype=np.uint8)
    for k in range(4):
        codes[:, :, k] = (packed >> (2 * k)) & 3
    return codes.reshape(m, 4 * b)[:, :iid_count]


def _resolve_index(index, count):
    if index is None:
        return np.arange(count, dtype=np.int64)
    idx = np.asarray(index).astype(np.int64).reshape(-1)
    idx = np.where(idx < 0, idx + count, idx)
    if idx.size and (idx.min() < 0 or idx.max() >= count):
        raise IndexError("index out of range for axis of size {0}".format(count))
    return idx


def decode(packed, iid_count, iid_index=None, sid_index=None, count_A1=False,
           dtype=np.float64, order="F"):
    """val[a, b] = decode(iid_index[a], sid_index[b]) in the requested dtype / order."""
    dtype = np.dtype(dtype)
    sid_count = packed.shape[0]
    ii = _resolve_index(iid_index, iid_count)
    si = _resolve_index(sid_index, sid_count)
    codes = unpack_codes(packed[si], iid_count)[:, ii]          # [n_s, n_i]
    lut = np.array(_LUT_A1 if count_A1 else _LUT_A2, dtype=np.float64)
    if dtype == np.int8:
        lut = np.where(np.isnan(lut), INT8_MISSING, lut)
    val = lut[codes].astype(dtype)                              # [n_s, n_i]
    if order == "A":
        order = "F"
    return np.asarray(val.T, order=order)                        # [n_i, n_s]


def beta_pdf(x, a, b):
    """Beta(a,b) density, same closed form SciPy evaluates (standardizer.py:204-205)."""
    x = np.asarray(x, dtype=np.float64)
    lnB = math.lgamma(a) + math.lgamma(b) - math.lgamma(a + b)
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = np.where(a == 1.0, 0.0, (a - 1.0) * np.log(x))
        t2 = np.where(b == 1.0, 0.0, (b - 1.0) * np.log1p(-x))
        out = np.exp(t1 + t2 - lnB)
    return out


def standardize(val, is_beta=False, a=np.nan, b=np.nan, use_stats=False, stats=None):
    """Restates ``_standardize_unit_python`` / ``_standardize_beta_python`` in float64.

    Returns (standardized float64 array [N, M], stats float64 [M, 2]).  The input is not
    modified.  NaN -> 0 afterwards; an SNC SNP (std == 0) gets std = inf and a zero column.
    """
    x = np.array(val, dtype=np.float64, order="F")
    miss = np.isnan(x)
    if use_stats:
        st = np.array(stats, dtype=np.float64)
        mean, std = st[:, 0].copy(), st[:, 1].copy()
    else:
        n_obs = (~miss).sum(0).astype(np.float64)
        with np.errstate(invalid="ignore", divide="ignore"):
            mean = np.where(miss, 0.0, x).sum(0) / n_obs
            dev = np.where(miss, 0.0, x - mean)
            std = np.sqrt((dev * dev).sum(0) / n_obs)
        std[std == 0.0] = np.inf
        st = np.stack([mean, std], axis=1)
    if is_beta:
        maf = mean / 2.0
        maf = np.where(maf > 0.5, 1.0 - maf, maf)
        factor = beta_pdf(maf, float(a), float(b))
        with np.errstate(invalid="ignore"):
            out = (x - mean) * factor
        out[:, np.isinf(std)] = 0.0          # SNC in training data -> 0 (standardizer.py:210-211)
    else:
        with np.errstate(invalid="ignore"):
            out = (x - mean) / std
    out[miss] = 0.0
    return out, st


def kernel(x):
    """K = X X^T in float64 (snpdata.py:203-206)."""
    x = np.asarray(x, dtype=np.float64)
    return x.dot(x.T)


def read_kernel(packed, iid_count, is_beta=False, a=np.nan, b=np.nan, count_A1=False,
                block_size=None, iid_index=None, sid_index=None):
    """Block loop of ``SnpReader._read_kernel`` (snpreader.py:623-668) in float64."""
    si = _resolve_index(sid_index, packed.shape[0])
    ii = _resolve_index(iid_index, iid_count)
    n = ii.size
    if block_size is None or si.size <= block_size or si.size <= n:
        block_size = max(si.size, 1)
    K = np.zeros((n, n), dtype=np.float64)
    stats = []
    for start in range(0, si.size, block_size):
        raw = decode(packed, iid_count, ii, si[start:start + block_size], count_A1, np.float64, "F")
        xs, st = standardize(raw, is_beta, a, b)
        stats.append(st)
        K += kernel(xs)
    return K, (np.concatenate(stats) if stats else np.zeros((0, 2)))


def read_cross_kernel(packed_r, iid_count_r, packed_c, iid_count_c, is_beta=False, a=np.nan, b=np.nan, count_A1_r=False,
                      count_A1_c=False, iid_index_r=None, iid_index_c=None, sid_index_r=None, sid_index_c=None, stats=None):
    """Train x test kernel in float64: standardize the row (train) side (``standardizer.py:135-211``; or apply ``stats``),
    apply ITS statistics to the column (test) side (``unittrained.py:47-70`` / ``betatrained.py:47-63``: use_stats=True),
    then ``train.val.dot(test.val.T)`` (the product of ``snpdata.py:203-206`` with two operands)."""
    xr = decode(packed_r, iid_count_r, _resolve_index(iid_index_r, iid_count_r), _resolve_index(sid_index_r, packed_r.shape[0]),
                count_A1_r, np.float64, "F")
    xc = decode(packed_c, iid_count_c, _resolve_index(iid_index_c, iid_count_c), _resolve_index(sid_index_c, packed_c.shape[0]),
                count_A1_c, np.float64, "F")
    if stats is None:
        xr, st = standardize(xr, is_beta, a, b)
    else:
        xr, st = standardize(xr, is_beta, a, b, use_stats=True, stats=stats)
    xc, _ = standardize(xc, is_beta, a, b, use_stats=True, stats=st)
    return xr.dot(xc.T), st


def sub_matrix(val, row_index, col_index, dtype=None, order="C"):
    """``out[i,j,...] = val[row[i], col[j], ...]`` (util/__init__.py:271-393)."""
    out = np.asarray(val)[np.asarray(row_index, dtype=np.int64)][:, np.asarray(col_index, dtype=np.int64)]
    return np.asarray(out, dtype=dtype or val.dtype, order=order)


# ----------------------------------------------------------------------------------
# synthetic packed genotypes: p_j ~ U(.05,.5), g ~ Binomial(2,p_j),
# missing iid-Bernoulli(r); per-chunk RNG streams so any SNP range is reproducible.
# ----------------------------------------------------------------------------------
SYNTH_CHUNK = 4096
_CODE_OF_DOSAGE_A2 = np.array([0, 2, 3], dtype=np.uint8)   # 0->00, 1->10, 2->11 ; missing -> 01


def synth_packed(iid_count, sid_start, sid_stop, missing_rate=0.0, seed=0):
    """Packed records [sid_stop-sid_start, ceil(N/4)] of the synthetic .bed (count_A1=False coding)."""
    b = bytes_per_snp(iid_count)
    out = np.zeros((sid_stop - sid_start, b), dtype=np.uint8)
    c0, c1 = sid_start // SYNTH_CHUNK, (sid_stop - 1) // SYNTH_CHUNK if sid_stop > sid_start else -1
    for c in range(c0, c1 + 1):
        rng = np.random.default_rng([seed, c])
        p = rng.uniform(0.05, 0.5, SYNTH_CHUNK)
        lo, hi = max(sid_start, c * SYNTH_CHUNK), min(sid_stop, (c + 1) * SYNTH_CHUNK)
        for j in range(lo, hi):
            rj = np.random.default_rng([seed, c, j - c * SYNTH_CHUNK])
            g = rj.binomial(2, p[j - c * SYNTH_CHUNK], iid_count)
            code = _CODE_OF_DOSAGE_A2[g]
            if missing_rate > 0:
                code = np.where(rj.random(iid_count) < missing_rate, np.uint8(1), code)
            pad = np.zeros(4 * b, dtype=np.uint8)
            pad[:iid_count] = code
            q = pad.reshape(b, 4)
            out[j - sid_start] = q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)
    return out


def write_bed(path, packed):
    with open(path, "wb") as f:
        f.write(BED_MAGIC)
        f.write(np.ascontiguousarray(packed).tobytes())
