// runtime.cu -- housekeeping entry points of libpst_b200.so: error string, launch counter, device facts.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>
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

// zeroed launch counter for the kernels that feed work through an atomic (k_syrk2's tile feed, the records of the F-order read kernels):
// a small per-thread, per-device ring, cleared in stream order before every launch
int next_counter(cudaStream_t st, int** out) {
    constexpr int kRing = 256;
    static thread_local int* d_ring = nullptr;
    static thread_local int ring_dev = -1;
    static thread_local unsigned seq = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail("cudaGetDevice failed");
    if (!d_ring || ring_dev != dev) {
        d_ring = nullptr;                                   // (a ring of another device is left to that device's teardown)
        cudaError_t e = cudaMalloc(&d_ring, kRing * sizeof(int));
        if (e != cudaSuccess) return fail("cudaMalloc(counter ring) -> %s", cudaGetErrorString(e));
        ring_dev = dev;
    }
    int* c = d_ring + (seq++ % kRing);
    cudaError_t e = cudaMemsetAsync(c, 0, sizeof(int), st);
    if (e != cudaSuccess) return fail("cudaMemsetAsync(counter) -> %s", cudaGetErrorString(e));
    *out = c;
    return 0;
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

// Occupancy probe for kernel experiments (not part of the reference-facing ABI): how many clusters of `cluster_size` CTAs with
// `threads` threads and `smem_bytes` of dynamic shared memory can be co-resident on the current device.
namespace pstb {
__global__ void k_cluster_probe(int* out) {
    extern __shared__ uint8_t probe_smem[];
    if (out && threadIdx.x == 0 && blockIdx.x == 0x7fffffff) *out = (int)probe_smem[0];
}
}  // namespace pstb

// NUMA placement for one-process-per-GPU hosts (include/pst_b200.h).  No libnuma in the image: /sys for the topology,
// sched_setaffinity + the raw set_mempolicy syscall for the binding.
extern "C" int pstb_numa_bind(int device) {
    char bdf[32] = {};
    if (cudaDeviceGetPCIBusId(bdf, (int)sizeof(bdf), device) != cudaSuccess) {
        cudaGetLastError();
        pstb::fail("pstb_numa_bind: no PCI bus id for device %d", device);
        return -2;
    }
    for (char* c = bdf; *c; ++c) *c = (char)tolower((unsigned char)*c);
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bdf);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    if (node < 0) return -1;
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f) return -1;
    char list[4096] = {};
    const bool got = fgets(list, sizeof(list), f) != nullptr;
    fclose(f);
    if (!got) return -1;
    cpu_set_t allowed, want;
    CPU_ZERO(&want);
    if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) CPU_ZERO(&allowed);
    int picked = 0;
    char* save = nullptr;
    for (char* tok = strtok_r(list, ",\n", &save); tok; tok = strtok_r(nullptr, ",\n", &save)) {       // "0-15,64-79"
        int a = 0, b = 0;
        const int n = sscanf(tok, "%d-%d", &a, &b);
        if (n < 1) continue;
        if (n == 1) b = a;
        for (int c = a; c <= b && c < CPU_SETSIZE; ++c)
            if (CPU_ISSET(c, &allowed)) { CPU_SET(c, &want); ++picked; }
    }
    if (picked == 0) return -1;                                  // the node's CPUs are not ours to use (cgroup / container mask)
    if (sched_setaffinity(0, sizeof(want), &want) != 0) {
        pstb::fail("pstb_numa_bind: sched_setaffinity to node %d failed", node);
        return -2;
    }
#ifdef SYS_set_mempolicy
    if (node < 64) {
        unsigned long mask = 1ul << node;
        syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, &mask, 65ul);     // best effort: first touch from the bound CPUs is local anyway
    }
#endif
    return node;
}

extern "C" int pstb_debug_max_active_clusters(int cluster_size, int threads, int smem_bytes) {
    if (cluster_size < 1 || threads < 1 || smem_bytes < 0) return -1;
    if (cudaFuncSetAttribute(pstb::k_cluster_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) return -2;
    if (cluster_size > 8 && cudaFuncSetAttribute(pstb::k_cluster_probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return -3;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(cluster_size * 64), 1, 1);
    cfg.blockDim = dim3((unsigned)threads, 1, 1);
    cfg.dynamicSmemBytes = (size_t)smem_bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster_size;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, pstb::k_cluster_probe, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return -4;
    }
    return n;
}
