// standardize.cu -- K2f (standardize an existing float matrix in place), gather (sub_matrix) and K0 (pack).
//
// K2f replaces bed_reader.standardize_f32 / standardize_f64 as called from
// pysnptools/standardizer/standardizer.py:109-121; semantics follow the reference's python twins
// (standardizer.py:135-163 Unit, :175-211 Beta): per SNP over non-NaN entries mean and population sd
// (two-pass), sd == 0 -> inf, (x-mean)/sd or (x-mean)*BetaPDF(maf), NaN -> 0.
// The gather replaces subset_f64_f64 / f32_f64 / f32_f32 (util/__init__.py:341-375), the packer replaces
// to_bed's write_f32/f64/i8 (bed.py:300-314).
#include "pstb_common.cuh"

namespace pstb {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum over the whole CTA; every thread receives the result.  `scratch` holds >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int k = 0; k < nw; ++k) t += scratch[k];
    return t;
}

// value = (x - mean) * scale with ONE reciprocal per SNP instead of an fp64 division per value (the transform was bound by the
// fp64 pipe, not by HBM): scale = 1/sd (Unit; sd = inf -> 0) or the Beta weight (0 for an SNC SNP).  The subtraction stays in
// fp64; float32 outputs multiply in float32 like the reference's float32 path does.
__device__ __forceinline__ double std_scale(int mode, double sd, double f) {
    if (mode == PSTB_STD_UNIT) return 1.0 / sd;
    return isinf(sd) ? 0.0 : f;
}
template <typename T>
__device__ __forceinline__ T std_apply(double x, double mean, double scale);
template <>
__device__ __forceinline__ double std_apply<double>(double x, double mean, double scale) { return (x - mean) * scale; }
template <>
__device__ __forceinline__ float std_apply<float>(double x, double mean, double scale) { return (float)(x - mean) * (float)scale; }

// ---- F order: one CTA per SNP column --------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024) k_std_f(T* val, long long n_iid, long long n_sid, int mode, double a, double b,
                                               double lnB, int apply, int use_stats, double* stats) {
    __shared__ double scratch[32];
    for (long long j = blockIdx.x; j < n_sid; j += gridDim.x) {
        T* col = val + j * n_iid;
        double mean, sd;
        if (use_stats) {
            mean = stats[2 * j];
            sd = stats[2 * j + 1];
        } else {
            double s = 0.0, c = 0.0;
            for (long long i = threadIdx.x; i < n_iid; i += blockDim.x) {
                double x = (double)col[i];
                if (x == x) { s += x; c += 1.0; }
            }
            s = block_sum(s, scratch);
            c = block_sum(c, scratch);
            mean = s / c;
            double ss = 0.0;
            for (long long i = threadIdx.x; i < n_iid; i += blockDim.x) {
                double x = (double)col[i];
                if (x == x) ss += (x - mean) * (x - mean);
            }
            ss = block_sum(ss, scratch);
            sd = sqrt(ss / c);
            if (sd == 0.0) sd = INFINITY;
            if (threadIdx.x == 0) { stats[2 * j] = mean; stats[2 * j + 1] = sd; }
        }
        if (apply) {
            const double scale = std_scale(mode, sd, (mode == PSTB_STD_BETA) ? beta_factor(mean, a, b, lnB) : 0.0);
            for (long long i = threadIdx.x; i < n_iid; i += blockDim.x) {
                double x = (double)col[i];
                col[i] = (x != x) ? (T)0 : std_apply<T>(x, mean, scale);
            }
        }
    }
}

// ---- F order, column staged in shared memory: every value crosses HBM once in each direction ----------------------------
// k_std_f above sweeps a column three times (sum, squared deviations, apply) and relies on L2 for the re-reads.  Here a CTA
// brings the column into shared memory with ONE TMA bulk copy (the next column's copy is in flight meanwhile when two fit),
// runs the two statistics passes and the transform on the shared copy, and writes it back with ONE bulk store
// (cp.async.bulk.global.shared::cta): no load / store instructions touch global memory at all.
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename T>
__global__ void __launch_bounds__(1024) k_std_f_staged(T* val, long long n_iid, long long n_sid, int mode, double a, double b,
                                                      double lnB, int apply, int use_stats, double* stats, unsigned col_bytes,
                                                      int nbuf, int* counter) {
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    __shared__ double scratch[64];
    __shared__ int s_claim;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_dyn);
    unsigned char* buf0 = smem_dyn + 128;
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](long long jj, int buf) {
        mbar_expect_tx(&bars[buf], col_bytes);
        bulk_g2s(buf0 + (size_t)buf * col_bytes, val + jj * n_iid, col_bytes, &bars[buf]);
    };
    // columns are claimed from the launch's counter (one ahead), not dealt out by a static stride: see ReadParams::counter
    auto claim = [&]() -> long long {
        __syncthreads();                                  // the previous claim has been read by everyone
        if (tid == 0) s_claim = atomicAdd(counter, 1);
        __syncthreads();
        return (long long)s_claim;
    };
    long long j = claim(), nj = 0;
    if (tid == 0 && j < n_sid) issue(j, 0);
    for (uint32_t it = 0; j < n_sid; j = nj, ++it) {
        const int cur = (nbuf == 2) ? (int)(it & 1u) : 0;
        const uint32_t parity = (nbuf == 2) ? ((it >> 1) & 1u) : (it & 1u);
        T* col = reinterpret_cast<T*>(buf0 + (size_t)cur * col_bytes);
        nj = claim();
        if (nbuf == 2 && tid == 0 && nj < n_sid) {
            bulk_store_wait_read();                       // the store of the column that lived in the other buffer has read it
            issue(nj, cur ^ 1);
        }
        mbar_wait(&bars[cur], parity);
        // 128-bit shared-memory accesses: 4 float32 / 2 float64 values per instruction (the scalar version was issue-bound at
        // 39 instructions per value, profiles/r1_late_kernels_full.txt); the column starts 16-byte aligned in shared memory
        constexpr int VW = 16 / sizeof(T);
        struct alignas(16) Vec { T v[VW]; };
        Vec* colv = reinterpret_cast<Vec*>(col);
        const long long nvec = n_iid / VW;
        double mean, sd;
        if (use_stats) {
            mean = stats[2 * j];
            sd = stats[2 * j + 1];
        } else {
            double s = 0.0;
            unsigned int cnt = 0;
            for (long long i = tid; i < nvec; i += blockDim.x) {
                const Vec x = colv[i];
#pragma unroll
                for (int k = 0; k < VW; ++k) {
                    const bool ok = x.v[k] == x.v[k];
                    s += ok ? (double)x.v[k] : 0.0;
                    cnt += ok ? 1u : 0u;
                }
            }
            for (long long i = nvec * VW + tid; i < n_iid; i += blockDim.x) {
                const T x = col[i];
                const bool ok = x == x;
                s += ok ? (double)x : 0.0;
                cnt += ok ? 1u : 0u;
            }
            // sum and count share one reduction: the count rides in the second half of the scratch array
            s = warp_sum(s);
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            const int w = tid >> 5, nw = (blockDim.x + 31) >> 5;
            __syncthreads();
            if ((tid & 31) == 0) { scratch[w] = s; scratch[32 + w] = (double)cnt; }
            __syncthreads();
            s = 0.0;
            double c = 0.0;
            for (int k = 0; k < nw; ++k) { s += scratch[k]; c += scratch[32 + k]; }
            mean = s / c;
            double ss = 0.0;
            for (long long i = tid; i < nvec; i += blockDim.x) {
                const Vec x = colv[i];
#pragma unroll
                for (int k = 0; k < VW; ++k) {
                    const double d = (x.v[k] == x.v[k]) ? (double)x.v[k] - mean : 0.0;
                    ss = fma(d, d, ss);
                }
            }
            for (long long i = nvec * VW + tid; i < n_iid; i += blockDim.x) {
                const T x = col[i];
                const double d = (x == x) ? (double)x - mean : 0.0;
                ss = fma(d, d, ss);
            }
            ss = block_sum(ss, scratch);
            sd = sqrt(ss / c);
            if (sd == 0.0) sd = INFINITY;
            if (tid == 0) { stats[2 * j] = mean; stats[2 * j + 1] = sd; }
        }
        if (apply) {
            const double scale = std_scale(mode, sd, (mode == PSTB_STD_BETA) ? beta_factor(mean, a, b, lnB) : 0.0);
            for (long long i = tid; i < nvec; i += blockDim.x) {
                Vec x = colv[i];
#pragma unroll
                for (int k = 0; k < VW; ++k) x.v[k] = (x.v[k] != x.v[k]) ? (T)0 : std_apply<T>((double)x.v[k], mean, scale);
                colv[i] = x;
            }
            for (long long i = nvec * VW + tid; i < n_iid; i += blockDim.x) {
                const T x = col[i];
                col[i] = (x != x) ? (T)0 : std_apply<T>((double)x, mean, scale);
            }
            fence_async_smem();                           // generic-proxy writes -> visible to the bulk store
            __syncthreads();
            if (tid == 0) bulk_s2g(val + j * n_iid, col, col_bytes);
        } else {
            __syncthreads();                              // every thread is done with the column before its buffer is refilled
        }
        if (nbuf == 1 && tid == 0 && nj < n_sid) {
            bulk_store_wait_read();
            issue(nj, 0);
        }
    }
    if (tid == 0) bulk_store_wait_read();                 // shared memory must outlive the last store's reads
}

// ---- C order: one sweep for the statistics, one for the transform ---------------------------------------------------------
// Thread <-> SNP column, so a CTA reads whole 1-2 KiB row pieces; the rows are cut into up to kMaxSplits blocks.  Every
// (block, column) pair yields (n, mean, M2) from ONE pass with sums shifted by the block's first valid value
// (s1 = sum(x - k), s2 = sum((x - k)^2): no cancellation as long as k is a sample of the column), and the blocks are merged
// in a fixed order with the pairwise update of Chan et al. -- deterministic, and equal to the two-pass nanmean / nanstd of
// the reference's python twin to ~1e-15.  (The first version swept DRAM three times and spent a 64-bit modulo per value.)
// work layout (doubles): [0,m) mean  [m,2m) scale  then per block b: [2m + 3mb, +m) n  [.. + m, +m) mean  [.. + 2m, +m) M2
constexpr int kMaxSplits = 12;

template <typename T>
__global__ void __launch_bounds__(256) k_colstats_c(const T* val, long long n_iid, long long n_sid, long long rows_per_block,
                                                    double* work) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_sid) return;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = min(n_iid, r0 + rows_per_block);
    const T* p = val + r0 * n_sid + j;
    double k = 0.0, s1 = 0.0, s2 = 0.0;
    unsigned int cnt = 0;
    bool have_k = false;
    long long i = r0;
    for (; i + 4 <= r1; i += 4, p += 4 * n_sid) {
        const T x0 = p[0], x1 = p[n_sid], x2 = p[2 * n_sid], x3 = p[3 * n_sid];        // four independent loads in flight
        const T xs[4] = {x0, x1, x2, x3};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double x = (double)xs[u];
            if (x == x) {
                if (!have_k) { k = x; have_k = true; }
                const double d = x - k;
                s1 += d;
                s2 = fma(d, d, s2);
                ++cnt;
            }
        }
    }
    for (; i < r1; ++i, p += n_sid) {
        const double x = (double)p[0];
        if (x == x) {
            if (!have_k) { k = x; have_k = true; }
            const double d = x - k;
            s1 += d;
            s2 = fma(d, d, s2);
            ++cnt;
        }
    }
    double* part = work + 2 * n_sid + (long long)blockIdx.y * 3 * n_sid;
    const double n = (double)cnt;
    part[j] = n;
    part[n_sid + j] = cnt ? k + s1 / n : 0.0;
    part[2 * n_sid + j] = cnt ? s2 - s1 * s1 / n : 0.0;
}

__global__ void k_finalize_c(long long n_sid, int splits, int mode, double a, double b, double lnB, int use_stats, double* stats,
                             double* work) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_sid) return;
    double mean, sd;
    if (use_stats) {
        mean = stats[2 * j];
        sd = stats[2 * j + 1];
    } else {
        double n = 0.0, m2 = 0.0;
        mean = 0.0;
        for (int s = 0; s < splits; ++s) {                       // fixed order: deterministic
            const double* part = work + 2 * n_sid + (long long)s * 3 * n_sid;
            const double nb = part[j], mb = part[n_sid + j], m2b = part[2 * n_sid + j];
            if (nb > 0.0) {
                const double nt = n + nb, delta = mb - mean;
                mean += delta * (nb / nt);
                m2 += m2b + delta * delta * (n * nb / nt);
                n = nt;
            }
        }
        if (n == 0.0) mean = NAN;                                // all-missing SNP: NaN statistics, as the python twin
        sd = sqrt(m2 / n);
        if (sd == 0.0) sd = INFINITY;
        stats[2 * j] = mean;
        stats[2 * j + 1] = sd;
    }
    work[j] = mean;
    work[n_sid + j] = std_scale(mode, sd, (mode == PSTB_STD_BETA) ? beta_factor(mean, a, b, lnB) : 0.0);
}

template <typename T>
__global__ void __launch_bounds__(256) k_apply_c(T* val, long long n_iid, long long n_sid, long long rows_per_block, const double* work) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_sid) return;
    const double mean = work[j], scale = work[n_sid + j];
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = min(n_iid, r0 + rows_per_block);
    T* p = val + r0 * n_sid + j;
    long long i = r0;
    for (; i + 4 <= r1; i += 4, p += 4 * n_sid) {
        const T x0 = p[0], x1 = p[n_sid], x2 = p[2 * n_sid], x3 = p[3 * n_sid];
        p[0] = (x0 != x0) ? (T)0 : std_apply<T>((double)x0, mean, scale);
        p[n_sid] = (x1 != x1) ? (T)0 : std_apply<T>((double)x1, mean, scale);
        p[2 * n_sid] = (x2 != x2) ? (T)0 : std_apply<T>((double)x2, mean, scale);
        p[3 * n_sid] = (x3 != x3) ? (T)0 : std_apply<T>((double)x3, mean, scale);
    }
    for (; i < r1; ++i, p += n_sid) {
        const T x = p[0];
        p[0] = (x != x) ? (T)0 : std_apply<T>((double)x, mean, scale);
    }
}

template <typename T>
static int standardize_impl(T* d_val, int order, int64_t n_iid, int64_t n_sid, int mode, double a, double b, double lnB,
                            int apply, int use_stats, double* d_stats, double* d_work, cudaStream_t st) {
    const int sms = sm_count_cached();
    if (order == PSTB_ORDER_F) {
        const unsigned long long col_bytes = (unsigned long long)n_iid * sizeof(T);
        const unsigned max_smem = 220u * 1024u;
        static const bool v1 = getenv("PSTB_STD_F_V1") && atoi(getenv("PSTB_STD_F_V1")) != 0;            // A/B runs
        if (!v1 && col_bytes >= 4096 && col_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(d_val) & 15u) == 0 && 128 + col_bytes <= max_smem) {
            // measured on B200 (scripts/sweep_k2f.sh): ONE buffer per CTA and as many CTAs as shared memory holds beat double
            // buffering inside fewer CTAs (float32 N = 10 000: 4.9 vs 3.8 TB/s) -- neighbouring CTAs overlap each other's load /
            // compute / store phases; threads so that the resident CTAs fill the SM's 2048 thread slots
            int nbuf = 1;
            const int fit = (int)(max_smem / (128 + col_bytes));
            int threads = fit >= 4 ? 256 : (fit >= 2 ? 512 : 1024);
            if (const char* e = getenv("PSTB_STD_NBUF")) { int v = atoi(e); if (v == 1 || (v == 2 && 128 + 2 * col_bytes <= max_smem)) nbuf = v; }   // tuning
            if (const char* e = getenv("PSTB_STD_THREADS")) { int v = atoi(e); if (v >= 64 && v <= 1024 && v % 32 == 0) threads = v; }
            const unsigned smem = 128u + (unsigned)nbuf * (unsigned)col_bytes;
            PSTB_CUDA(cudaFuncSetAttribute(k_std_f_staged<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int ctas_per_sm = 1;
            PSTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_std_f_staged<T>, threads, smem));
            if (ctas_per_sm < 1) ctas_per_sm = 1;
            long long grid = (long long)sms * ctas_per_sm;
            if (grid > n_sid) grid = n_sid;
            int* counter = nullptr;
            if (next_counter(st, &counter)) return 1;
            k_std_f_staged<T><<<(unsigned)grid, threads, smem, st>>>(d_val, n_iid, n_sid, mode, a, b, lnB, apply, use_stats, d_stats,
                                                                     (unsigned)col_bytes, nbuf, counter);
            PSTB_AFTER_LAUNCH("k_std_f_staged");
            return 0;
        }
        // columns that do not fit shared memory: three sweeps, the second and third out of L2 -- so only as many columns in
        // flight as L2 holds (64 MB budget of the 126 MB)
        long long per_sm = (long long)((64ull << 20) / ((unsigned long long)sms * (col_bytes ? col_bytes : 1)));
        per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
        const int threads = per_sm <= 2 ? 1024 : 256;
        long long grid = n_sid < (long long)sms * per_sm ? n_sid : (long long)sms * per_sm;
        k_std_f<T><<<(unsigned)grid, threads, 0, st>>>(d_val, n_iid, n_sid, mode, a, b, lnB, apply, use_stats, d_stats);
        PSTB_AFTER_LAUNCH("k_std_f");
        return 0;
    }
    if (!d_work) return fail("C-order standardize needs d_work (pstb_standardize_work_bytes)");
    const unsigned gx = (unsigned)((n_sid + 255) / 256);
    long long splits = ((long long)sms * 8 + gx - 1) / gx;
    if (splits < 1) splits = 1;
    if (splits > kMaxSplits) splits = kMaxSplits;
    long long rows_per_block = (n_iid + splits - 1) / splits;
    if (rows_per_block < 4) rows_per_block = 4;
    unsigned gy = (unsigned)((n_iid + rows_per_block - 1) / rows_per_block);
    if (gy < 1) gy = 1;
    if (!use_stats) {
        k_colstats_c<T><<<dim3(gx, gy), 256, 0, st>>>(d_val, n_iid, n_sid, rows_per_block, d_work);
        PSTB_AFTER_LAUNCH("k_colstats_c");
    }
    k_finalize_c<<<(unsigned)((n_sid + 255) / 256), 256, 0, st>>>(n_sid, (int)gy, mode, a, b, lnB, use_stats, d_stats, d_work);
    PSTB_AFTER_LAUNCH("k_finalize_c");
    if (apply) {
        // the transform wants more CTAs than the statistics pass has row blocks: cut the rows finer
        long long rows_apply = (n_iid + 63) / 64;
        if (rows_apply < 16) rows_apply = 16;
        const unsigned gya = (unsigned)((n_iid + rows_apply - 1) / rows_apply);
        k_apply_c<T><<<dim3(gx, gya < 1 ? 1 : gya), 256, 0, st>>>(d_val, n_iid, n_sid, rows_apply, d_work);
        PSTB_AFTER_LAUNCH("k_apply_c");
    }
    return 0;
}

// ---- gather -------------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) k_subset(const TI* in, long long si, long long sj, long long sk, Axis rows, Axis cols,
                                                long long v, TO* out, int order_out, long long n_in, long long m_in) {
    const long long n = rows.n, m = cols.n, total = n * m * v;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        long long i, j, k;
        if (order_out == PSTB_ORDER_C) { k = e % v; j = (e / v) % m; i = e / (v * m); }
        else { i = e % n; j = (e / n) % m; k = e / (n * m); }
        long long r = rows.at(i), c = cols.at(j);
        r = r < 0 ? 0 : (r >= n_in ? n_in - 1 : r);
        c = c < 0 ? 0 : (c >= m_in ? m_in - 1 : c);
        out[e] = (TO)in[r * si + c * sj + k * sk];
    }
}

// ---- pack -----------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ int code_of(T x, int count_a1, int* bad);
template <>
__device__ __forceinline__ int code_of<int8_t>(int8_t x, int count_a1, int* bad) {
    if (x == 0) return count_a1 ? 3 : 0;
    if (x == 1) return 2;
    if (x == 2) return count_a1 ? 0 : 3;
    if (x != -127) *bad = 1;
    return 1;
}
template <typename T>
__device__ __forceinline__ int code_of(T x, int count_a1, int* bad) {
    if (x == (T)0) return count_a1 ? 3 : 0;
    if (x == (T)1) return 2;
    if (x == (T)2) return count_a1 ? 0 : 3;
    if (x == x) *bad = 1;
    return 1;
}

// One thread packs 16 genotypes of one SNP into a 32-bit word.  F order (individuals contiguous) reads them with 128-bit
// loads when the column is 16-byte aligned; every other layout walks the strides.
template <typename T>
__global__ void __launch_bounds__(256) k_pack(const T* val, long long si, long long sj, long long n_iid, long long n_sid,
                                              int count_a1, uint8_t* packed, long long ld, int32_t* d_bad) {
    const long long rec = (n_iid + 3) / 4, words = (n_iid + 15) / 16, total = words * n_sid;
    const bool word_store = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(packed) & 3u) == 0);
    int bad = 0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        // adjacent threads read adjacent memory: next 16-genotype word of the same SNP (F order) or next SNP (C order)
        const long long j = (sj == 1) ? e % n_sid : e / words, w = (sj == 1) ? e / n_sid : e % words, i0 = w * 16;
        const T* src = val + i0 * si + j * sj;
        uint32_t word = 0;
        if (si == 1 && i0 + 16 <= n_iid && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
            T v[16];
            const uint4* s4 = reinterpret_cast<const uint4*>(src);
            uint4* d4 = reinterpret_cast<uint4*>(v);
#pragma unroll
            for (int k = 0; k < (int)sizeof(T); ++k) d4[k] = __ldg(s4 + k);
#pragma unroll
            for (int t = 0; t < 16; ++t) word |= (uint32_t)code_of<T>(v[t], count_a1, &bad) << (2 * t);
        } else {
            for (int t = 0; t < 16; ++t)
                if (i0 + t < n_iid) word |= (uint32_t)code_of<T>(src[t * si], count_a1, &bad) << (2 * t);
        }
        uint8_t* dst = packed + j * ld + 4 * w;
        if (word_store && 4 * w + 4 <= rec) {
            *reinterpret_cast<uint32_t*>(dst) = word;
        } else {
            for (int k = 0; k < 4; ++k)
                if (4 * w + k < rec) dst[k] = (uint8_t)(word >> (8 * k));
        }
    }
    if (bad && d_bad) *d_bad = 1;
}

}  // namespace pstb

using namespace pstb;

extern "C" int64_t pstb_standardize_work_bytes(int64_t n_sid) { return (int64_t)((2 + 3 * kMaxSplits) * (n_sid > 0 ? n_sid : 1) * sizeof(double)); }

extern "C" int pstb_standardize(void* d_val, int dtype, int order, int64_t n_iid, int64_t n_sid, int mode, double a, double b,
                                int apply_in_place, int use_stats, double* d_stats, void* d_work, void* stream) {
    if (n_iid < 0 || n_sid < 0) return fail("negative shape");
    if (mode != PSTB_STD_UNIT && mode != PSTB_STD_BETA) return fail("mode must be PSTB_STD_UNIT or PSTB_STD_BETA");
    if (mode == PSTB_STD_BETA && !(a > 0.0 && b > 0.0)) return fail("Beta parameters must be positive");
    if (order != PSTB_ORDER_F && order != PSTB_ORDER_C) return fail("bad order");
    if (n_sid == 0) return 0;
    if (!d_stats) return fail("d_stats is NULL");
    if (n_iid > 0 && !d_val) return fail("d_val is NULL");
    const double lnB = (mode == PSTB_STD_BETA) ? lgamma(a) + lgamma(b) - lgamma(a + b) : 0.0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == PSTB_F32)
        return standardize_impl<float>((float*)d_val, order, n_iid, n_sid, mode, a, b, lnB, apply_in_place, use_stats, d_stats,
                                       (double*)d_work, st);
    if (dtype == PSTB_F64)
        return standardize_impl<double>((double*)d_val, order, n_iid, n_sid, mode, a, b, lnB, apply_in_place, use_stats, d_stats,
                                        (double*)d_work, st);
    return fail("standardize needs float32 or float64");
}

extern "C" int pstb_subset(const void* d_in, int dtype_in, int order_in, int64_t n_in, int64_t m_in, int64_t v, pstb_axis rows,
                           pstb_axis cols, void* d_out, int dtype_out, int order_out, void* stream) {
    if (n_in < 0 || m_in < 0 || v < 0 || rows.n < 0 || cols.n < 0) return fail("negative shape");
    const long long total = (long long)rows.n * cols.n * v;
    if (total == 0) return 0;
    if (!d_in || !d_out) return fail("NULL array");
    for (int ax = 0; ax < 2; ++ax) {
        const pstb_axis& s = ax ? cols : rows;
        const int64_t cnt = ax ? m_in : n_in;
        if (!s.idx) {
            int64_t last = s.start + (s.n - 1) * s.step;
            if (s.start < 0 || s.start >= cnt || last < 0 || last >= cnt) return fail("subset selection out of range");
        }
    }
    long long si, sj, sk;
    if (order_in == PSTB_ORDER_C) { si = m_in * v; sj = v; sk = 1; }
    else if (order_in == PSTB_ORDER_F) { si = 1; sj = n_in; sk = n_in * m_in; }
    else return fail("bad order_in");
    if (order_out != PSTB_ORDER_C && order_out != PSTB_ORDER_F) return fail("bad order_out");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    long long grid = (total + 255) / 256;
    const long long cap = (long long)sm_count_cached() * 16;
    if (grid > cap) grid = cap;
    Axis r = to_axis(rows), c = to_axis(cols);
    if (dtype_in == PSTB_F64 && dtype_out == PSTB_F64)
        k_subset<double, double><<<(unsigned)grid, 256, 0, st>>>((const double*)d_in, si, sj, sk, r, c, v, (double*)d_out, order_out, n_in, m_in);
    else if (dtype_in == PSTB_F32 && dtype_out == PSTB_F64)
        k_subset<float, double><<<(unsigned)grid, 256, 0, st>>>((const float*)d_in, si, sj, sk, r, c, v, (double*)d_out, order_out, n_in, m_in);
    else if (dtype_in == PSTB_F32 && dtype_out == PSTB_F32)
        k_subset<float, float><<<(unsigned)grid, 256, 0, st>>>((const float*)d_in, si, sj, sk, r, c, v, (float*)d_out, order_out, n_in, m_in);
    else if (dtype_in == PSTB_F64 && dtype_out == PSTB_F32)
        k_subset<double, float><<<(unsigned)grid, 256, 0, st>>>((const double*)d_in, si, sj, sk, r, c, v, (float*)d_out, order_out, n_in, m_in);
    else
        return fail("subset supports float32 / float64");
    PSTB_AFTER_LAUNCH("k_subset");
    return 0;
}

extern "C" int pstb_pack(const void* d_val, int dtype, int order, int64_t n_iid, int64_t n_sid, int count_a1, uint8_t* d_packed,
                         int64_t ld, int32_t* d_bad, void* stream) {
    if (n_iid < 0 || n_sid < 0) return fail("negative shape");
    const long long rec = (n_iid + 3) / 4;
    if (ld < rec) return fail("ld smaller than ceil(iid_count/4)");
    if (rec * n_sid == 0) return 0;
    if (!d_val || !d_packed) return fail("NULL array");
    long long si, sj;
    if (order == PSTB_ORDER_C) { si = n_sid; sj = 1; }
    else if (order == PSTB_ORDER_F) { si = 1; sj = n_iid; }
    else return fail("bad order");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    long long grid = (((n_iid + 15) / 16) * n_sid + 255) / 256;
    const long long cap = (long long)sm_count_cached() * 16;
    if (grid > cap) grid = cap;
    if (dtype == PSTB_F32) k_pack<float><<<(unsigned)grid, 256, 0, st>>>((const float*)d_val, si, sj, n_iid, n_sid, count_a1, d_packed, ld, d_bad);
    else if (dtype == PSTB_F64) k_pack<double><<<(unsigned)grid, 256, 0, st>>>((const double*)d_val, si, sj, n_iid, n_sid, count_a1, d_packed, ld, d_bad);
    else if (dtype == PSTB_I8) k_pack<int8_t><<<(unsigned)grid, 256, 0, st>>>((const int8_t*)d_val, si, sj, n_iid, n_sid, count_a1, d_packed, ld, d_bad);
    else return fail("bad dtype");
    PSTB_AFTER_LAUNCH("k_pack");
    return 0;
}
