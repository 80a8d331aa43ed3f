// pstb_common.cuh -- shared helpers of libpst_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/pst_b200.h"

namespace pstb {

int fail(const char* fmt, ...);           // records the thread-local message, returns 1
void count_launch(int n = 1);             // feeds pstb_launch_count()
int sm_count_cached();
int next_counter(cudaStream_t st, int** out);   // zeroed device int for an atomic work feed (ring, cleared in stream order)

// selection along one axis as the kernels see it
struct Axis {
    const uint32_t* idx;
    long long start, step, n;
    __host__ __device__ inline long long at(long long k) const {
#ifdef __CUDA_ARCH__
        return idx ? (long long)__ldg(idx + k) : start + k * step;
#else
        return idx ? (long long)idx[k] : start + k * step;
#endif
    }
};
inline Axis to_axis(const pstb_axis& a) { return Axis{a.idx, (long long)a.start, (long long)a.step, (long long)a.n}; }

// per-SNP factors ------------------------------------------------------------------------------
// Unit:  v = (g - mean) / sd                      (standardizer.py:158-161)
// Beta:  v = (g - mean) * BetaPDF(maf; a, b)      (standardizer.py:198-211); lnB = lgamma(a)+lgamma(b)-lgamma(a+b)
__device__ __forceinline__ double beta_factor(double mean, double a, double b, double lnB) {
    double maf = mean * 0.5;
    if (maf > 0.5) maf = 1.0 - maf;
    if (!(maf >= 0.0 && maf <= 1.0)) return (maf != maf) ? maf : 0.0;
    double t1 = (a == 1.0) ? 0.0 : (a - 1.0) * log(maf);
    double t2 = (b == 1.0) ? 0.0 : (b - 1.0) * log1p(-maf);
    return exp(t1 + t2 - lnB);
}

// value of dosage g (0,1,2) under (mode, mean, sd); `f` is beta_factor() for Beta (unused for Unit)
__device__ __forceinline__ double std_value(int mode, double g, double mean, double sd, double f) {
    if (mode == PSTB_STD_UNIT) return (g - mean) / sd;
    if (isinf(sd)) return 0.0;
    return (g - mean) * f;
}

// mean / population sd from exact dosage counts (two-pass formula of nanstd, evaluated on counts)
__device__ __forceinline__ void stats_from_counts(long long n0, long long n1, long long n2, double& mean, double& sd) {
    double n = (double)(n0 + n1 + n2);
    mean = ((double)n1 + 2.0 * (double)n2) / n;            // n == 0 -> NaN, as the python twin
    double d0 = 0.0 - mean, d1 = 1.0 - mean, d2 = 2.0 - mean;
    double ss = (double)n0 * d0 * d0 + (double)n1 * d1 * d1 + (double)n2 * d2 * d2;
    sd = sqrt(ss / n);
    if (sd == 0.0) sd = INFINITY;                            // SNC -> inf (standardizer.py:154)
}

// ---- mbarrier / bulk-copy (TMA 1-D) primitives -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// bounded wait: a lost arrival traps after ~4 s instead of hanging the GPU box
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const unsigned long long t0 = global_ns();
    for (uint32_t spin = 1;; ++spin) {
        if (mbar_try_wait(bar, parity)) return;
        if ((spin & 1023u) == 0 && global_ns() - t0 > 4000000000ull) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <typename T>
__device__ __forceinline__ T from_double(double v);
template <>
__device__ __forceinline__ float from_double<float>(double v) { return (float)v; }
template <>
__device__ __forceinline__ double from_double<double>(double v) { return v; }
template <>
__device__ __forceinline__ int8_t from_double<int8_t>(double v) { return (int8_t)v; }

// float32 -> float32 / float64 copy of `total` contiguous entries with a scalar factor (syrk.cu)
int convert_range(const float* d_K, long long total, void* d_out, int dtype, double scale, void* stream);

// one slice of a streamed pstb_snp_kernel (syrk.cu): phase bit 0 = first slice, bit 1 = last slice; the same d_work throughout;
// total_sid = SNPs of the whole kernel (the low-term mode "auto" decides on it)
int snp_kernel_slice(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count, pstb_axis iid, pstb_axis sid,
                     int count_a1, int mode, double a, double b, int use_stats, double* d_stats, float* d_K, int accumulate,
                     void* d_work, int64_t work_bytes, int64_t chunk, int low_term, void* stream, int phase, int64_t total_sid);

// one slice of a streamed float64 kernel (syrk_f64.cu): decode + standardize into a float64 panel, fp64 FMA SYRK, no mirror
int snp_kernel_f64_slice(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count, pstb_axis iid, pstb_axis sid,
                         int count_a1, int mode, double a, double b, int use_stats, double* d_stats, double* d_K, int accumulate,
                         void* d_work, int64_t work_bytes, int64_t chunk, void* stream);

}  // namespace pstb

#define PSTB_CUDA(x)                                                                   \
    do {                                                                               \
        cudaError_t e__ = (x);                                                         \
        if (e__ != cudaSuccess) return pstb::fail("%s -> %s", #x, cudaGetErrorString(e__)); \
    } while (0)

#define PSTB_AFTER_LAUNCH(name)                                                                 \
    do {                                                                                        \
        pstb::count_launch();                                                                   \
        cudaError_t e__ = cudaGetLastError();                                                   \
        if (e__ != cudaSuccess) return pstb::fail("launch of %s -> %s", name, cudaGetErrorString(e__)); \
    } while (0)
