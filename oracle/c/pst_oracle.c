/* CPU oracle in plain C -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A restatement of the reference's CPU path for decode + standardize, used (a) to cross-check
 * oracle/bed_oracle.py and (b) as the timed multi-threaded "port" CPU baseline in bench.py
 * (cpu_baseline / --impl reference).  Nothing under pysnptools_b200/ links or calls it.
 *
 * Follows:
 *   decode      - bed_reader.open_bed(...).read as called from pysnptools/snpreader/bed.py:337-343;
 *                 PLINK .bed layout: per SNP ceil(N/4) bytes, LSB pair first,
 *                 00->0, 01->missing, 10->1, 11->2 (count_A1=False) / 00->2,...,11->0 (count_A1=True).
 *                 One thread per SNP column, like the reference's rayon loop over SNPs.
 *   standardize - pysnptools/standardizer/standardizer.py:135-163 (Unit) and :175-211 (Beta):
 *                 per SNP over non-NaN: mean, population std (two-pass), std==0 -> inf,
 *                 (x-mean)/std or (x-mean)*BetaPDF(maf), NaN -> 0, SNC column -> 0.
 * Build: make -C oracle   (gcc -O3 -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static double beta_pdf(double x, double a, double b) {
    double lnB = lgamma(a) + lgamma(b) - lgamma(a + b);
    double t1 = (a == 1.0) ? 0.0 : (a - 1.0) * log(x);
    double t2 = (b == 1.0) ? 0.0 : (b - 1.0) * log1p(-x);
    return exp(t1 + t2 - lnB);
}

int pst_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* out is F-order [n_iid, n_sid] (order_c == 0) or C-order (order_c == 1). Index vectors may be NULL (= all). */
#define DEFINE_DECODE(NAME, T, MISSING)                                                                   \
    int NAME(const uint8_t* packed, int64_t bytes_per_snp, int64_t iid_count, int64_t sid_count,          \
             const int64_t* iid_idx, int64_t n_iid, const int64_t* sid_idx, int64_t n_sid, int count_a1,   \
             int order_c, T* out, int num_threads) {                                                      \
        const T lut_a2[4] = {(T)0, (T)(MISSING), (T)1, (T)2};                                             \
        const T lut_a1[4] = {(T)2, (T)(MISSING), (T)1, (T)0};                                             \
        const T* lut = count_a1 ? lut_a1 : lut_a2;                                                        \
        int bad = 0;                                                                                      \
        (void)num_threads;                                                                                \
        _Pragma("omp parallel for schedule(static) num_threads(num_threads) reduction(|:bad)")           \
        for (int64_t b = 0; b < n_sid; ++b) {                                                             \
            int64_t j = sid_idx ? sid_idx[b] : b;                                                         \
            if (j < 0 || j >= sid_count) { bad = 1; continue; }                                           \
            const uint8_t* row = packed + j * bytes_per_snp;                                              \
            for (int64_t a = 0; a < n_iid; ++a) {                                                         \
                int64_t i = iid_idx ? iid_idx[a] : a;                                                     \
                if (i < 0 || i >= iid_count) { bad = 1; continue; }                                       \
                T v = lut[(row[i >> 2] >> (2 * (i & 3))) & 3];                                            \
                if (order_c) out[a * n_sid + b] = v; else out[a + b * n_iid] = v;                         \
            }                                                                                             \
        }                                                                                                 \
        return bad;                                                                                       \
    }

DEFINE_DECODE(pst_oracle_decode_f32, float, NAN)
DEFINE_DECODE(pst_oracle_decode_f64, double, NAN)
DEFINE_DECODE(pst_oracle_decode_i8, int8_t, -127)

/* In-place standardize of val[n_iid, n_sid]; stats is [n_sid][2] double (C order). */
#define DEFINE_STD(NAME, T)                                                                               \
    int NAME(T* val, int64_t n_iid, int64_t n_sid, int order_c, int is_beta, double a, double b,          \
             int use_stats, double* stats, int num_threads) {                                             \
        int64_t si = order_c ? n_sid : 1, sj = order_c ? 1 : n_iid;                                       \
        (void)num_threads;                                                                                \
        _Pragma("omp parallel for schedule(static) num_threads(num_threads)")                            \
        for (int64_t j = 0; j < n_sid; ++j) {                                                             \
            T* col = val + j * sj;                                                                        \
            double mean, sd;                                                                              \
            if (use_stats) { mean = stats[2 * j]; sd = stats[2 * j + 1]; }                                \
            else {                                                                                        \
                double s = 0.0, n = 0.0;                                                                  \
                for (int64_t i = 0; i < n_iid; ++i) { double x = col[i * si]; if (x == x) { s += x; n += 1.0; } } \
                mean = s / n;                                                                             \
                double ss = 0.0;                                                                          \
                for (int64_t i = 0; i < n_iid; ++i) { double x = col[i * si]; if (x == x) { ss += (x - mean) * (x - mean); } } \
                sd = sqrt(ss / n);                                                                        \
                if (sd == 0.0) sd = INFINITY;                                                             \
                stats[2 * j] = mean; stats[2 * j + 1] = sd;                                               \
            }                                                                                             \
            if (is_beta) {                                                                                \
                double maf = mean / 2.0; if (maf > 0.5) maf = 1.0 - maf;                                  \
                double f = beta_pdf(maf, a, b); int snc = isinf(sd);                                      \
                for (int64_t i = 0; i < n_iid; ++i) { double x = col[i * si];                             \
                    col[i * si] = (x != x || snc) ? (T)0 : (T)((x - mean) * f); }                         \
            } else {                                                                                      \
                for (int64_t i = 0; i < n_iid; ++i) { double x = col[i * si];                             \
                    col[i * si] = (x != x) ? (T)0 : (T)((x - mean) / sd); }                               \
            }                                                                                             \
        }                                                                                                 \
        return 0;                                                                                         \
    }

DEFINE_STD(pst_oracle_standardize_f32, float)
DEFINE_STD(pst_oracle_standardize_f64, double)
