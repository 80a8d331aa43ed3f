// runtime.cu -- housekeeping entry points of libpst_b200.so: error string, launch counter, device facts.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include "pstb_common.cuh"

namespace pstb {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count_cached() {
    static thread_local int dev_cached = -1, sms = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != dev_cached) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        sms = v;
        dev_cached = dev;
    }
    return sms;
}

}  // namespace pstb

extern "C" int pstb_version(void) { return 100; }

extern "C" const char* pstb_last_error(void) { return pstb::g_err; }

extern "C" int pstb_sm_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return 0;
    }
    return pstb::sm_count_cached();
}

extern "C" int64_t pstb_packed_ld(int64_t iid_count) {
    int64_t rec = (iid_count + 3) / 4;
    return (rec + 15) / 16 * 16;
}

extern "C" int64_t pstb_launch_count(void) { return (int64_t)pstb::g_launches.load(); }

extern "C" void* pstb_host_alloc(int64_t bytes) {
    void* p = nullptr;
    if (bytes <= 0) bytes = 1;
    cudaError_t e = cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        pstb::fail("cudaHostAlloc(%lld) -> %s", (long long)bytes, cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}

extern "C" int pstb_host_free(void* p) {
    if (!p) return 0;
    cudaError_t e = cudaFreeHost(p);
    if (e != cudaSuccess) return pstb::fail("cudaFreeHost -> %s", cudaGetErrorString(e));
    return 0;
}
