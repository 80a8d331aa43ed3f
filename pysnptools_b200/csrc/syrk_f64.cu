// syrk_f64.cu -- K3 in float64: K = X X^T with fp64 FMA arithmetic on the CUDA cores (sm_100a).
//
// The reference computes its kernels in the dtype the caller asks for -- `val.dot(val.T)` (snpdata.py:203-206) is a DGEMM for the
// default dtype=float64 and an SGEMM for float32 -- and its own unit tests compare float64 kernels to 10 decimals
// (kernelreader/test.py:48-50, :189; test.py:535-553).  The tensor-core path (syrk.cu) accumulates in fp32 (<= 1e-5 relative
// Frobenius error, the north_star gate): right for float32 requests, not for those.  This file is the float64 twin:
//   per chunk of SNPs:  fused decode + exact-count statistics + standardize into a float64 panel [n, chunk] (decode.cu, the values
//                       the reference's standardize_f64 produces to ~1e-16), then
//   k_dsyrk:            K_lower (+)= panel panel^T, 128 x 128 tiles of the lower triangle, 8 x 8 accumulators per thread, 16-deep
//                       k-steps staged in shared memory with the next step prefetched into registers.  One partial sum per chunk
//                       is formed in registers and then added to K, i.e. a two-level summation.
// B200 has full-rate fp64 FMA units (40 TFLOP/s nominal), so this path is ~100 x the reference's float64 CPU rate while matching
// it to ~1e-13 relative.
#include <cstdlib>
#include "pstb_common.cuh"

namespace pstb {
int read_impl_ex(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count, pstb_axis iid, pstb_axis sid,
                 int count_a1, int mode, double a, double b, int use_stats, double* d_stats, void* d_out, int dtype, int order,
                 void* stream, unsigned int* d_miss_flag);

namespace {

constexpr int DT = 128;   // tile edge
constexpr int DK = 16;    // k-step

// X(i, j) = X[i * si + j * sj]: n rows (individuals), kc columns (SNPs).  ROW_FAST: si == 1 (F order), else sj == 1 (C order).
template <bool ROW_FAST>
__global__ void __launch_bounds__(256, 1) k_dsyrk(const double* __restrict__ X, long long si, long long sj, long long n, long long kc,
                                                  double* __restrict__ K, long long ldk, int accumulate) {
    __shared__ __align__(16) double As[DK][DT];
    __shared__ __align__(16) double Bs[DK][DT];
    // lower-triangular tile (I, J), J <= I, from the linear block index
    const long long t = blockIdx.x;
    long long I = (long long)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while (I * (I + 1) / 2 > t) --I;
    while ((I + 1) * (I + 2) / 2 <= t) ++I;
    const long long J = t - I * (I + 1) / 2;
    const long long i0 = I * DT, k0 = J * DT;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    // global -> register staging: 2 x 8 values per thread and k-step
    int li[8], lk[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int lin = tid + 256 * e;                    // 0 .. 2047
        if (ROW_FAST) { li[e] = lin & (DT - 1); lk[e] = lin >> 7; }       // consecutive threads -> consecutive individuals
        else { lk[e] = lin & (DK - 1); li[e] = lin >> 4; }                // consecutive threads -> consecutive SNPs
    }
    double ra[8], rb[8];
    auto fetch = [&](long long kk) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const long long j = kk + lk[e];
            const long long ia = i0 + li[e], ib = k0 + li[e];
            ra[e] = (j < kc && ia < n) ? X[ia * si + j * sj] : 0.0;
            rb[e] = (j < kc && ib < n) ? X[ib * si + j * sj] : 0.0;
        }
    };
    double acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.0;
    fetch(0);
    for (long long kk = 0; kk < kc; kk += DK) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { As[lk[e]][li[e]] = ra[e]; Bs[lk[e]][li[e]] = rb[e]; }
        __syncthreads();
        if (kk + DK < kc) fetch(kk + DK);                  // next k-step in flight while this one is multiplied
#pragma unroll
        for (int k = 0; k < DK; ++k) {
            double a[8], b[8];
            // rows ty*8 .. ty*8+7 (64 contiguous bytes, two distinct addresses per warp: broadcast); columns tx + 16 c (conflict-free)
            const double2* ap = reinterpret_cast<const double2*>(&As[k][ty * 8]);
#pragma unroll
            for (int r = 0; r < 4; ++r) { const double2 v = ap[r]; a[2 * r] = v.x; a[2 * r + 1] = v.y; }
#pragma unroll
            for (int c = 0; c < 8; ++c) b[c] = Bs[k][tx + 16 * c];
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const long long i = i0 + ty * 8 + r;
        if (i >= n) continue;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const long long k = k0 + tx + 16 * c;
            if (k >= n || k > i) continue;                 // lower triangle only (diagonal tiles compute both halves)
            double* dst = K + i * ldk + k;
            *dst = accumulate ? *dst + acc[r][c] : acc[r][c];
        }
    }
}

__global__ void __launch_bounds__(256) k_mirror_f64(double* K, long long n, long long ldk) {
    __shared__ double tile[32][33];
    const long long bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const long long i = bi * 32 + r, j = bj * 32 + tx;
        tile[r][tx] = (i < n && j < n) ? K[i * ldk + j] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const long long j = bj * 32 + r, i = bi * 32 + tx;
        if (i < n && j < n && j < i) K[j * ldk + i] = tile[tx][r];
    }
}

int launch_dsyrk(const double* X, int order, long long n, long long kc, double* K, long long ldk, int accumulate, cudaStream_t st) {
    const long long T = (n + DT - 1) / DT, tiles = T * (T + 1) / 2;
    if (tiles < 1) return 0;
    if (tiles > 0x7fffffffLL) return fail("too many tiles");
    if (order == PSTB_ORDER_F) k_dsyrk<true><<<(unsigned)tiles, 256, 0, st>>>(X, 1, n, n, kc, K, ldk, accumulate);
    else k_dsyrk<false><<<(unsigned)tiles, 256, 0, st>>>(X, kc, 1, n, kc, K, ldk, accumulate);
    PSTB_AFTER_LAUNCH("k_dsyrk");
    return 0;
}

}  // namespace

// one slice of a streamed float64 kernel (pstb_snp_kernel_host_f64): no mirror, K accumulated when `accumulate`
int snp_kernel_f64_slice(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count, pstb_axis iid, pstb_axis sid,
                         int count_a1, int mode, double a, double b, int use_stats, double* d_stats, double* d_K, int accumulate,
                         void* d_work, int64_t work_bytes, int64_t chunk, void* stream) {
    if (mode != PSTB_STD_UNIT && mode != PSTB_STD_BETA) return fail("kernel needs PSTB_STD_UNIT or PSTB_STD_BETA");
    if (mode == PSTB_STD_BETA && !(a > 0.0 && b > 0.0)) return fail("Beta parameters must be positive");
    if (iid.n < 0 || sid.n < 0) return fail("negative selection length");
    if (iid.n == 0) return 0;
    if (!d_K) return fail("d_K is NULL");
    if (sid.n > 0 && !d_stats) return fail("d_stats is NULL");
    if (chunk < 1) return fail("chunk must be positive");
    if (work_bytes < pstb_kernel_f64_workspace_bytes(iid.n, chunk) || !d_work) return fail("workspace too small (pstb_kernel_f64_workspace_bytes)");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long n = iid.n;
    if (sid.n == 0) {
        if (!accumulate) PSTB_CUDA(cudaMemsetAsync(d_K, 0, (size_t)n * n * sizeof(double), st));
        return 0;
    }
    double* panel = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(d_work) + 255) & ~(uintptr_t)255);
    for (long long c0 = 0; c0 < sid.n; c0 += chunk) {
        const long long ns = (c0 + chunk <= sid.n) ? chunk : sid.n - c0;
        pstb_axis sub = sid;
        sub.n = ns;
        if (sid.idx) sub.idx = sid.idx + c0; else sub.start = sid.start + c0 * sid.step;
        // decode + statistics + standardize, float64, F order: panel[i + j * n] -- missing -> 0, SNC -> 0 (standardizer.py:145-163)
        int rc = read_impl_ex(d_packed, ld, iid_count, sid_count, iid, sub, count_a1, mode, a, b, use_stats, d_stats + 2 * c0, panel, PSTB_F64,
                              PSTB_ORDER_F, stream, nullptr);
        if (rc) return rc;
        rc = launch_dsyrk(panel, PSTB_ORDER_F, n, ns, d_K, n, (accumulate || c0 > 0) ? 1 : 0, st);
        if (rc) return rc;
    }
    return 0;
}

}  // namespace pstb

using namespace pstb;

extern "C" int64_t pstb_kernel_f64_workspace_bytes(int64_t n_iid, int64_t chunk) {
    if (n_iid < 1) n_iid = 1;
    if (chunk < 1) chunk = 1;
    return (int64_t)((size_t)n_iid * (size_t)chunk * sizeof(double) + 512);
}

extern "C" int pstb_mirror_lower_f64(double* d_K, int64_t n, int64_t ldk, void* stream) {
    if (n <= 0) return 0;
    if (!d_K) return fail("NULL pointer");
    const unsigned g = (unsigned)((n + 31) / 32);
    k_mirror_f64<<<dim3(g, g), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_K, n, ldk);
    PSTB_AFTER_LAUNCH("k_mirror_f64");
    return 0;
}

extern "C" int pstb_snp_kernel_f64(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count, pstb_axis iid, pstb_axis sid,
                                   int count_a1, int mode, double a, double b, int use_stats, double* d_stats, double* d_K, int accumulate,
                                   int mirror, void* d_work, int64_t work_bytes, int64_t chunk, void* stream) {
    int rc = snp_kernel_f64_slice(d_packed, ld, iid_count, sid_count, iid, sid, count_a1, mode, a, b, use_stats, d_stats, d_K, accumulate, d_work,
                                  work_bytes, chunk, stream);
    if (rc) return rc;
    if (mirror && iid.n > 0) return pstb_mirror_lower_f64(d_K, iid.n, iid.n, stream);
    return 0;
}

extern "C" int pstb_float_kernel_f64(const double* d_val, int order, int64_t n_iid, int64_t n_sid, double* d_K, int accumulate, int mirror,
                                     void* stream) {
    if (n_iid < 0 || n_sid < 0) return fail("negative shape");
    if (n_iid == 0) return 0;
    if (!d_K) return fail("d_K is NULL");
    if (order != PSTB_ORDER_C && order != PSTB_ORDER_F) return fail("bad order");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (n_sid == 0) {
        if (!accumulate) PSTB_CUDA(cudaMemsetAsync(d_K, 0, (size_t)n_iid * n_iid * sizeof(double), st));
        return 0;
    }
    if (!d_val) return fail("d_val is NULL");
    int rc = launch_dsyrk(d_val, order, n_iid, n_sid, d_K, n_iid, accumulate ? 1 : 0, st);
    if (rc) return rc;
    if (mirror) return pstb_mirror_lower_f64(d_K, n_iid, n_iid, stream);
    return 0;
}
