// host_api.cu -- host-buffer entry points: what a bed_reader-style binding calls with NumPy arrays.
//
// pstb_read_host streams the selected SNP records through the GPU in chunks on two CUDA streams:
// H2D of packed records (2 bits / genotype), the fused decode(+standardize) kernel, and D2H of the float
// output overlap, so the call runs at the speed of the device->host link.  Pinned caller buffers
// (pstb_host_alloc) are copied directly; pageable ones go through internal pinned staging.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include "pstb_common.cuh"

namespace pstb {
int read_impl(const uint8_t* d_packed, int64_t ld, int64_t iid_count, int64_t sid_count, pstb_axis iid, pstb_axis sid,
              int count_a1, int mode, double a, double b, int use_stats, double* d_stats, void* d_out, int dtype, int order,
              void* stream);

namespace {

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    bool host = false;
    int ensure(size_t n) {
        if (n <= cap) return 0;
        release();
        if (n < 256) n = 256;
        cudaError_t e = host ? cudaHostAlloc(&p, n, cudaHostAllocDefault) : cudaMalloc(&p, n);
        if (e != cudaSuccess) {
            p = nullptr;
            cap = 0;
            return fail("%s(%zu bytes) -> %s", host ? "cudaHostAlloc" : "cudaMalloc", n, cudaGetErrorString(e));
        }
        cap = n;
        return 0;
    }
    void release() {
        if (p) {
            if (host) cudaFreeHost(p); else cudaFree(p);
        }
        p = nullptr;
        cap = 0;
    }
};

constexpr int kSlots = 4;     // pipeline depth of the host-buffer read: H2D, repack, kernel and D2H of different chunks overlap
constexpr int kMaxSlots = 16; // ... and with a pageable destination: chunks in flight between the GPU and the host copy workers

struct HostCtx {
    int device = -1;
    cudaStream_t s[kSlots] = {};
    cudaEvent_t done[kMaxSlots] = {}, in_done[kMaxSlots] = {};
    Buf d_packed[kMaxSlots], d_tight[kMaxSlots], d_out[kMaxSlots], d_stats, d_idx, d_work, d_K, h_in[kMaxSlots], h_out[kMaxSlots];
    Buf d_tiles, d_tail_work, d_tail_packed, d_u;           // pstb_snp_kernel_host with the copy-out of K overlapped (compact tiles, band-major tail)
    cudaStream_t hp = nullptr;                             // high priority: the expansion of finished bands runs between the SYRK launches
    cudaEvent_t copied[2] = {}, used[2] = {};
    HostCtx() {
        for (int k = 0; k < kMaxSlots; ++k) { h_in[k].host = true; h_out[k].host = true; }
    }
    void release_buffers() {
        for (int k = 0; k < kMaxSlots; ++k) { d_packed[k].release(); d_tight[k].release(); d_out[k].release(); h_in[k].release(); h_out[k].release(); }
        d_stats.release(); d_idx.release(); d_work.release(); d_K.release();
        d_tiles.release(); d_tail_work.release(); d_tail_packed.release(); d_u.release();
    }
    int init() {
        int dev = 0;
        PSTB_CUDA(cudaGetDevice(&dev));
        if (dev == device) return 0;
        for (int k = 0; k < kMaxSlots; ++k) {
            d_packed[k].release(); d_tight[k].release(); d_out[k].release(); h_in[k].release(); h_out[k].release();
            if (done[k]) cudaEventDestroy(done[k]);
            if (in_done[k]) cudaEventDestroy(in_done[k]);
            PSTB_CUDA(cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming));
            PSTB_CUDA(cudaEventCreateWithFlags(&in_done[k], cudaEventDisableTiming));
        }
        for (int k = 0; k < kSlots; ++k) {
            if (s[k]) cudaStreamDestroy(s[k]);
            PSTB_CUDA(cudaStreamCreateWithFlags(&s[k], cudaStreamNonBlocking));
        }
        d_stats.release(); d_idx.release(); d_work.release(); d_K.release();
        d_tiles.release(); d_tail_work.release(); d_tail_packed.release(); d_u.release();
        if (hp) cudaStreamDestroy(hp);
        {
            int lo_prio = 0, hi_prio = 0;
            PSTB_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
            PSTB_CUDA(cudaStreamCreateWithPriority(&hp, cudaStreamNonBlocking, hi_prio));
        }
        for (int k = 0; k < 2; ++k) {
            if (copied[k]) cudaEventDestroy(copied[k]);
            if (used[k]) cudaEventDestroy(used[k]);
            PSTB_CUDA(cudaEventCreateWithFlags(&copied[k], cudaEventDisableTiming));
            PSTB_CUDA(cudaEventCreateWithFlags(&used[k], cudaEventDisableTiming));
        }
        device = dev;
        return 0;
    }
};

HostCtx& ctx() {
    static thread_local HostCtx c;
    return c;
}

bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

// pageable destinations: copy out of the pinned ring with several host threads (a single memcpy stream tops out near 10 GB/s and
// first-touch page faults of a fresh NumPy array are serial otherwise).  Measured on a 16-core B200 host (scripts/probe_pagefault.py):
// a fresh array fills at 35 GB/s with 8 threads and 51 GB/s with 16 (107 GB/s once its pages exist), so the default is all cores but
// two (the caller's thread keeps enqueueing GPU work, the CUDA driver has threads of its own), at most 16.
int host_copy_threads() {
    if (const char* e = getenv("PSTB_HOST_COPY_THREADS")) {          // experiments
        const int v = atoi(e);
        if (v >= 1 && v <= 256) return v;
    }
    static int n = [] {
        unsigned hc = std::thread::hardware_concurrency();
        int v = hc ? (int)hc - 2 : 4;
        return v < 2 ? 2 : (v > 16 ? 16 : v);
    }();
    return n;
}

// A persistent pool of copy threads per calling thread.  Round 1 spawned fresh std::threads for every 64 MiB chunk; their creation is
// serial in the caller (~0.1 ms each), so 8 threads delivered 28 GB/s inside the pipeline against 77 GB/s for the same memcpy alone.
class CopyPool {
  public:
    ~CopyPool() { stop(); }
    template <typename F>
    void run(size_t count, size_t min_per_thread, F&& body, int max_threads = 0) {
        int nt = host_copy_threads();
        if (max_threads > 0 && nt > max_threads) nt = max_threads;
        if (count < 2 * min_per_thread) nt = 1;
        if ((size_t)nt > count / (min_per_thread ? min_per_thread : 1)) nt = (int)(count / (min_per_thread ? min_per_thread : 1));
        if (nt <= 1) { body((size_t)0, count); return; }
        ensure(nt - 1);
        const size_t per = (count + nt - 1) / nt;
        std::function<void(size_t, size_t)> fn = std::ref(body);
        {
            std::unique_lock<std::mutex> lk(m_);
            job_ = &fn;
            count_ = count;
            per_ = per;
            active_ = nt - 1;
            pending_ = nt - 1;
            ++generation_;
        }
        cv_.notify_all();
        body((size_t)0, per < count ? per : count);                 // the caller takes the first range
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [&] { return pending_ == 0; });
        job_ = nullptr;
    }

  private:
    void ensure(int workers) {
        while ((int)th_.size() < workers) {
            const int id = (int)th_.size();
            th_.emplace_back([this, id] { loop(id); });
        }
    }
    void loop(int id) {
        long long seen = 0;
        for (;;) {
            std::function<void(size_t, size_t)>* job = nullptr;
            size_t lo = 0, hi = 0;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                if (id >= active_) continue;                        // this job uses fewer workers
                job = job_;
                lo = (size_t)(id + 1) * per_;
                hi = lo + per_ < count_ ? lo + per_ : count_;
            }
            if (job && lo < hi) (*job)(lo, hi);
            std::unique_lock<std::mutex> lk(m_);
            if (--pending_ == 0) done_.notify_one();
        }
    }
    void stop() {
        {
            std::unique_lock<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : th_) t.join();
        th_.clear();
    }
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    std::function<void(size_t, size_t)>* job_ = nullptr;
    size_t count_ = 0, per_ = 0;
    int active_ = 0, pending_ = 0;
    long long generation_ = 0;
    bool stop_ = false;
};

// Workers that each take a WHOLE staged chunk of a pageable destination: wait for its D2H copy (CUDA event), copy it out of the pinned
// ring, free the slot.  No barrier per chunk -- with the fork-join pool above every chunk waited for its slowest thread (a huge-page
// fault here, a late wake-up there) and the pipeline delivered 25-28 GB/s into a fresh array where the same threads, left alone, fill
// one at 38-39 GB/s next to the DMA stream (scripts/probe_pagefault.py).
class TaskPool {
  public:
    ~TaskPool() { stop(); }
    void ensure(int n) {
        while ((int)th_.size() < n) th_.emplace_back([this] { loop(); });
    }
    void submit(std::function<void()> f) {
        {
            std::unique_lock<std::mutex> lk(m_);
            q_.push_back(std::move(f));
        }
        cv_.notify_one();
    }
    void submit_front(std::function<void()> f) {                    // ahead of the queued tasks (input staging: the GPU is waiting for it)
        {
            std::unique_lock<std::mutex> lk(m_);
            q_.insert(q_.begin(), std::move(f));
        }
        cv_.notify_one();
    }

  private:
    void loop() {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                f = std::move(q_.front());
                q_.erase(q_.begin());
            }
            f();
        }
    }
    void stop() {
        {
            std::unique_lock<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : th_) t.join();
        th_.clear();
    }
    std::vector<std::thread> th_;
    std::vector<std::function<void()>> q_;
    std::mutex m_;
    std::condition_variable cv_;
    bool stop_ = false;
};

TaskPool& task_pool() {
    static thread_local TaskPool pool;
    return pool;
}

CopyPool& copy_pool() {
    static thread_local CopyPool pool;
    return pool;
}

template <typename F>
void parallel_ranges(size_t count, size_t min_per_thread, F&& body, int max_threads = 0) {
    copy_pool().run(count, min_per_thread, body, max_threads);
}

// Threads for staging a chunk's packed input (pageable source -> pinned ring).  The packed records are 1/16 (float32) or 1/32 (float64)
// of the output bytes, so ONE thread stages a chunk well inside the time its output needs on PCIe -- and more threads are worse than
// useless: waking a fork-join team for every chunk slowed the whole pipeline down (cfg2, pageable in, pinned out, scripts/prof_api_read2.py:
// 0.83 s with 1 thread = the pinned-input time, 0.93 s with 2, 1.02 s with 4, 1.07 s with 8-14).  int8 output (input = 1/4 of it) gets 4.
int stage_threads(size_t in_bytes, size_t out_bytes) {
    if (getenv("PSTB_HOST_COPY_THREADS")) return host_copy_threads();     // experiments
    if (out_bytes == 0) return host_copy_threads();
    const size_t r = (16 * in_bytes + out_bytes / 2) / out_bytes;
    const int cap = host_copy_threads();
    return r < 1 ? 1 : (r > (size_t)cap ? cap : (int)r);
}

size_t esize_of(int dtype) { return dtype == PSTB_F64 ? 8 : (dtype == PSTB_F32 ? 4 : 1); }

// host index vector -> device axis (arithmetic progressions stay implicit)
int make_axis(const int64_t* h_idx, int64_t n, int64_t count, const char* name, Buf& d_idx, cudaStream_t st, pstb_axis* out,
              std::vector<uint32_t>& scratch) {
    out->idx = nullptr;
    out->start = 0;
    out->step = 1;
    out->n = n;
    if (!h_idx) {
        out->n = count;
        return 0;
    }
    for (int64_t k = 0; k < n; ++k)
        if (h_idx[k] < 0 || h_idx[k] >= count) return fail("%s index %lld out of range [0, %lld)", name, (long long)h_idx[k], (long long)count);
    if (n <= 1) {
        out->start = n ? h_idx[0] : 0;
        return 0;
    }
    const int64_t step = h_idx[1] - h_idx[0];
    bool arith = step != 0;
    for (int64_t k = 2; k < n && arith; ++k) arith = (h_idx[k] - h_idx[k - 1]) == step;
    if (arith) {
        out->start = h_idx[0];
        out->step = step;
        return 0;
    }
    scratch.resize((size_t)n);
    for (int64_t k = 0; k < n; ++k) scratch[(size_t)k] = (uint32_t)h_idx[k];
    if (d_idx.ensure((size_t)n * sizeof(uint32_t))) return 1;
    PSTB_CUDA(cudaMemcpyAsync(d_idx.p, scratch.data(), (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    PSTB_CUDA(cudaStreamSynchronize(st));
    out->idx = (const uint32_t*)d_idx.p;
    return 0;
}

}  // namespace
}  // namespace pstb

using namespace pstb;

// d_acc[i] += d_x[i]: the deferred rank-one vectors of the calls of one overlapped kernel
__global__ void k_add_f64(double* acc, const double* x, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) acc[i] += x[i];
}


// Host -> device upload that is COMPLETE when it returns: cudaMemcpyAsync on one of the context's (non-blocking) streams followed by a
// stream synchronize.  A blocking cudaMemcpy on the legacy default stream may return once a pageable source has been staged, before the
// DMA lands, and non-blocking streams do not synchronise with the legacy stream (round-1 advisor finding).
static int upload_sync(void* d_dst, const void* h_src, size_t bytes, cudaStream_t st) {
    if (!bytes) return 0;
    PSTB_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
    PSTB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int pstb_read_host(const uint8_t* h_packed, int64_t iid_count, int64_t sid_count, const int64_t* h_iid_idx,
                              int64_t n_iid, const int64_t* h_sid_idx, int64_t n_sid, int count_a1, int mode, double a, double b,
                              int use_stats, double* h_stats, void* h_out, int dtype, int order) {
    if (iid_count < 0 || sid_count < 0) return fail("negative iid_count / sid_count");
    if (!h_iid_idx) n_iid = iid_count;
    if (!h_sid_idx) n_sid = sid_count;
    if (n_iid < 0 || n_sid < 0) return fail("negative selection length");
    if (iid_count > 0xfffffff0LL) return fail("iid_count too large");
    if (mode != PSTB_STD_NONE && !h_stats) return fail("h_stats is NULL");
    if (n_sid > 0)
        for (int64_t k = 0; h_sid_idx && k < n_sid; ++k)
            if (h_sid_idx[k] < 0 || h_sid_idx[k] >= sid_count)
                return fail("sid index %lld out of range [0, %lld)", (long long)h_sid_idx[k], (long long)sid_count);
    HostCtx& c = ctx();
    if (c.init()) return 1;
    std::vector<uint32_t> scratch;
    pstb_axis iid_ax;
    if (make_axis(h_iid_idx, n_iid, iid_count, "iid", c.d_idx, c.s[0], &iid_ax, scratch)) return 1;
    if (n_iid == 0 || n_sid == 0) return 0;
    if (!h_packed || !h_out) return fail("NULL host buffer");

    const size_t es = esize_of(dtype);
    const int64_t rec = (iid_count + 3) / 4;
    const int64_t ld = pstb_packed_ld(iid_count);
    const size_t col_bytes = (size_t)n_iid * es;
    const bool packed_pinned = is_pinned(h_packed), out_pinned = is_pinned(h_out);
    static const size_t target_mb = [] { const char* e = getenv("PSTB_HOST_CHUNK_MB"); int v = e ? atoi(e) : 0; return (size_t)((v >= 1 && v <= 4096) ? v : 0); }();
    // output bytes per pipeline chunk (tuning knob: PSTB_HOST_CHUNK_MB): 64 MiB straight into a pinned destination; 32 MiB pieces, each
    // drained by one host worker, into a pageable one
    const size_t target = (target_mb ? target_mb : (out_pinned ? 64 : 32)) << 20;
    int64_t chunk = (int64_t)(target / (col_bytes > (size_t)ld ? col_bytes : (size_t)ld));
    if (chunk < 1) chunk = 1;
    if (order == PSTB_ORDER_C && chunk < 64) chunk = 64;   // keep the strided D2H rows at >= 256 bytes
    if (chunk > n_sid) chunk = n_sid;
    // slots in flight: 4 for a pinned destination (the DMA is the only consumer); up to 16 for a pageable one, so that ~14 host workers
    // can each be copying a chunk while the GPU fills the next ones
    int nslots = kSlots;
    if (!out_pinned) {
        nslots = kMaxSlots;
        if (const char* e = getenv("PSTB_HOST_SLOTS")) { const int v = atoi(e); if (v >= 2 && v <= kMaxSlots) nslots = v; }
    }

    const bool any_stats = mode != PSTB_STD_NONE;
    if (any_stats) {
        if (c.d_stats.ensure((size_t)n_sid * 2 * sizeof(double))) return 1;
        if (use_stats && upload_sync(c.d_stats.p, h_stats, (size_t)n_sid * 2 * sizeof(double), c.s[0])) return 1;
    }
    for (int k = 0; k < nslots; ++k) {
        if (c.d_packed[k].ensure((size_t)chunk * ld) || c.d_out[k].ensure((size_t)chunk * col_bytes)) return 1;
        if (!out_pinned && c.h_out[k].ensure((size_t)chunk * col_bytes)) return 1;
    }
    // pageable destination: every staged chunk is handed to a worker of the calling thread's task pool
    std::mutex slot_m;
    std::condition_variable slot_cv;
    bool slot_busy[kMaxSlots] = {};
    std::atomic<int> worker_rc{0};
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (!out_pinned) task_pool().ensure(host_copy_threads());

    struct Pending { int64_t b0 = 0, ns = 0; bool active = false; } pend[kMaxSlots];
    const bool trace = getenv("PSTB_HOST_TRACE") != nullptr;
    double t_wait = 0.0, t_copy = 0.0, t_enq = 0.0, t_stage = 0.0;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double, std::milli>(now() - t0).count(); };
    // the copy of one staged chunk into the pageable destination (run by a pool worker)
    auto drain_chunk = [&](int slot, int64_t b0, int64_t ns) {
        const char* src = (const char*)c.h_out[slot].p;
        if (order == PSTB_ORDER_F) {
            memcpy((char*)h_out + (size_t)b0 * col_bytes, src, (size_t)ns * col_bytes);
        } else {
            for (size_t i = 0; i < (size_t)n_iid; ++i)
                memcpy((char*)h_out + (i * (size_t)n_sid + (size_t)b0) * es, src + i * (size_t)ns * es, (size_t)ns * es);
        }
    };
    auto finish = [&](int slot) -> int {
        if (!pend[slot].active) return 0;
        const auto tw = now();
        if (!out_pinned) {                                          // wait for the worker that drains this slot
            std::unique_lock<std::mutex> lk(slot_m);
            slot_cv.wait(lk, [&] { return !slot_busy[slot]; });
            t_wait += ms_since(tw);
            pend[slot].active = false;
            return worker_rc.load() ? fail("a host copy worker failed (CUDA error %d)", worker_rc.load()) : 0;
        }
        PSTB_CUDA(cudaEventSynchronize(c.done[slot]));
        t_wait += ms_since(tw);
        pend[slot].active = false;
        if (out_pinned) return 0;
        const auto tc = now();
        struct CopyTimer { double& acc; std::chrono::steady_clock::time_point t0; ~CopyTimer() { acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); } } copy_timer{t_copy, tc};
        const int64_t b0 = pend[slot].b0, ns = pend[slot].ns;
        const char* src = (const char*)c.h_out[slot].p;
        if (order == PSTB_ORDER_F) {
            char* dst = (char*)h_out + (size_t)b0 * col_bytes;
            parallel_ranges((size_t)ns * col_bytes, (size_t)1 << 20, [&](size_t lo, size_t hi) { memcpy(dst + lo, src + lo, hi - lo); });
        } else {
            parallel_ranges((size_t)n_iid, 256, [&](size_t lo, size_t hi) {
                for (size_t i = lo; i < hi; ++i)
                    memcpy((char*)h_out + (i * (size_t)n_sid + (size_t)b0) * es, src + i * (size_t)ns * es, (size_t)ns * es);
            });
        }
        return 0;
    };

    int rc = 0;
    int64_t nchunks = (n_sid + chunk - 1) / chunk;
    // a failure inside the pipelined loop leaves through the drain below (every slot finished, device synchronised): no async copy
    // into the caller's buffers may still be in flight when the error is returned
#define PSTB_CUDA_BREAK(x)                                                                       \
    {                                                                                            \
        cudaError_t e__ = (x);                                                                   \
        if (e__ != cudaSuccess) { rc = pstb::fail("%s -> %s", #x, cudaGetErrorString(e__)); break; } \
    }
    // Input staging (records in pageable memory, or a scattered SNP selection: rows -> the slot's pinned ring buffer at pitch ld).  The
    // ring buffer of a slot is free as soon as the H2D copy of the chunk that used it last is done (event in_done), long before that
    // chunk's output has left -- so staging never has to wait for finish(slot):
    //  * pinned destination: the caller stages chunk ch itself, one thread, BEFORE it waits for the slot (it would only be waiting);
    //  * pageable destination: all cores are busy draining output, the caller alone gets 2.6 GB/s out of the contended memory system
    //    and became the critical path (0.95 s for cfg2's 2.5 GB) -- the pool workers stage up to kLook chunks ahead instead, their
    //    tasks queued in front of the drain tasks.
    auto chunk_of = [&](int64_t ch2, int64_t& b0, int64_t& ns) {
        b0 = ch2 * chunk;
        ns = (b0 + chunk <= n_sid) ? chunk : n_sid - b0;
    };
    auto is_staged = [&](int64_t b0, int64_t ns) {
        if (!packed_pinned) return true;
        for (int64_t k = 1; h_sid_idx && k < ns; ++k)
            if (h_sid_idx[b0 + k] != h_sid_idx[b0] + k) return true;
        return false;
    };
    auto stage_rows = [&](char* stage, int64_t b0, size_t lo, size_t hi) {
        for (size_t k = lo; k < hi; ++k) {
            const int64_t j = h_sid_idx ? h_sid_idx[b0 + (int64_t)k] : b0 + (int64_t)k;
            memcpy(stage + k * (size_t)ld, h_packed + (size_t)j * rec, (size_t)rec);
        }
    };
    bool in_recorded[kMaxSlots] = {}, stage_ready[kMaxSlots] = {};
    int stage_inflight = 0;                                       // staging tasks not finished yet (guarded by slot_m)
    const int64_t kLook = nslots - 1 < 4 ? nslots - 1 : 4;
    int64_t next_stage = 0;
    for (int64_t ch = 0; ch < nchunks && !rc; ++ch) {
        const int slot = (int)(ch % nslots);
        int64_t b0, ns;
        chunk_of(ch, b0, ns);
        cudaStream_t st = c.s[slot % kSlots];
        // ---- input records ----
        const int64_t j0 = h_sid_idx ? h_sid_idx[b0] : b0;
        const bool staged_in = is_staged(b0, ns);
        if (!out_pinned) {
            const auto ts = now();
            for (; next_stage <= ch + kLook && next_stage < nchunks && !rc; ++next_stage) {
                int64_t b2, n2;
                chunk_of(next_stage, b2, n2);
                if (!is_staged(b2, n2)) continue;
                const int s2 = (int)(next_stage % nslots);
                if (c.h_in[s2].ensure((size_t)chunk * ld)) { rc = 1; break; }
                {
                    std::unique_lock<std::mutex> lk(slot_m);
                    stage_ready[s2] = false;
                    ++stage_inflight;
                }
                char* stage = (char*)c.h_in[s2].p;
                cudaEvent_t ev = in_recorded[s2] ? c.in_done[s2] : nullptr;
                task_pool().submit_front([&, s2, b2, n2, stage, ev, cur_dev] {
                    static thread_local int dev_set = -1;
                    if (dev_set != cur_dev) { cudaSetDevice(cur_dev); dev_set = cur_dev; }
                    const cudaError_t e = ev ? cudaEventSynchronize(ev) : cudaSuccess;     // the H2D copy that read this buffer last
                    if (e != cudaSuccess) worker_rc.store((int)e); else stage_rows(stage, b2, 0, (size_t)n2);
                    {
                        std::unique_lock<std::mutex> lk(slot_m);
                        stage_ready[s2] = true;
                        --stage_inflight;
                    }
                    slot_cv.notify_all();
                });
            }
            if (rc) break;
            if (staged_in) {
                std::unique_lock<std::mutex> lk(slot_m);
                slot_cv.wait(lk, [&] { return stage_ready[slot]; });
            }
            t_stage += ms_since(ts);
        } else if (staged_in) {
            const auto ts = now();
            if (c.h_in[slot].ensure((size_t)chunk * ld)) { rc = 1; break; }
            if (in_recorded[slot]) PSTB_CUDA_BREAK(cudaEventSynchronize(c.in_done[slot]));
            char* stage = (char*)c.h_in[slot].p;
            parallel_ranges((size_t)ns, 64, [&](size_t lo, size_t hi) { stage_rows(stage, b0, lo, hi); },
                            stage_threads((size_t)ns * rec, (size_t)ns * col_bytes));
            t_stage += ms_since(ts);
        }
        if ((rc = finish(slot))) break;
        const auto te = now();
        struct EnqTimer { double& acc; std::chrono::steady_clock::time_point t0; ~EnqTimer() { acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); } } enq_timer{t_enq, te};
        if (!staged_in) {
            // one contiguous DMA over PCIe (rows of ceil(N/4) bytes make a slow 2-D copy), then re-pitch to ld on the device
            if (ld == rec) {
                PSTB_CUDA_BREAK(cudaMemcpyAsync(c.d_packed[slot].p, h_packed + (size_t)j0 * rec, (size_t)ns * rec, cudaMemcpyHostToDevice, st));
            } else {
                if (c.d_tight[slot].ensure((size_t)chunk * rec)) { rc = 1; break; }
                PSTB_CUDA_BREAK(cudaMemcpyAsync(c.d_tight[slot].p, h_packed + (size_t)j0 * rec, (size_t)ns * rec, cudaMemcpyHostToDevice, st));
                PSTB_CUDA_BREAK(cudaMemcpy2DAsync(c.d_packed[slot].p, (size_t)ld, c.d_tight[slot].p, (size_t)rec, (size_t)rec, (size_t)ns,
                                            cudaMemcpyDeviceToDevice, st));
            }
        } else {
            PSTB_CUDA_BREAK(cudaMemcpyAsync(c.d_packed[slot].p, c.h_in[slot].p, (size_t)ns * ld, cudaMemcpyHostToDevice, st));
            PSTB_CUDA_BREAK(cudaEventRecord(c.in_done[slot], st));
            in_recorded[slot] = true;
        }
        // ---- kernel ----
        pstb_axis sid_ax{nullptr, 0, 1, ns};
        double* d_st = any_stats ? (double*)c.d_stats.p + 2 * b0 : nullptr;
        rc = read_impl((const uint8_t*)c.d_packed[slot].p, ld, iid_count, ns, iid_ax, sid_ax, count_a1, mode, a, b, use_stats, d_st,
                       c.d_out[slot].p, dtype, order, st);
        if (rc) break;
        // ---- output ----
        if (order == PSTB_ORDER_F) {
            void* dst = out_pinned ? (void*)((char*)h_out + (size_t)b0 * col_bytes) : c.h_out[slot].p;
            PSTB_CUDA_BREAK(cudaMemcpyAsync(dst, c.d_out[slot].p, (size_t)ns * col_bytes, cudaMemcpyDeviceToHost, st));
        } else if (out_pinned) {
            PSTB_CUDA_BREAK(cudaMemcpy2DAsync((char*)h_out + (size_t)b0 * es, (size_t)n_sid * es, c.d_out[slot].p, (size_t)ns * es,
                                        (size_t)ns * es, (size_t)n_iid, cudaMemcpyDeviceToHost, st));
        } else {
            PSTB_CUDA_BREAK(cudaMemcpyAsync(c.h_out[slot].p, c.d_out[slot].p, (size_t)ns * col_bytes, cudaMemcpyDeviceToHost, st));
        }
        PSTB_CUDA_BREAK(cudaEventRecord(c.done[slot], st));
        pend[slot].b0 = b0;
        pend[slot].ns = ns;
        pend[slot].active = true;
        if (!out_pinned) {
            {
                std::unique_lock<std::mutex> lk(slot_m);
                slot_busy[slot] = true;
            }
            cudaEvent_t ev = c.done[slot];
            task_pool().submit([&, slot, b0, ns, ev, cur_dev] {
                static thread_local int dev_set = -1;
                if (dev_set != cur_dev) { cudaSetDevice(cur_dev); dev_set = cur_dev; }
                const cudaError_t e = cudaEventSynchronize(ev);
                if (e != cudaSuccess) worker_rc.store((int)e); else drain_chunk(slot, b0, ns);
                {
                    std::unique_lock<std::mutex> lk(slot_m);
                    slot_busy[slot] = false;
                }
                slot_cv.notify_all();
            });
        }
    }
#undef PSTB_CUDA_BREAK
    {
        std::unique_lock<std::mutex> lk(slot_m);                    // (an error exit can leave staging tasks behind: they use this frame)
        slot_cv.wait(lk, [&] { return stage_inflight == 0; });
    }
    for (int k = 0; k < kMaxSlots; ++k) {
        int r2 = finish(k);
        if (!rc) rc = r2;
    }
    if (trace)
        fprintf(stderr, "[pstb_read_host] %lld chunks of %lld SNPs, %d slots, %d copy workers: staging input %.1f ms, waiting for slots %.1f ms, enqueue %.1f ms\n",
                (long long)nchunks, (long long)chunk, nslots, host_copy_threads(), t_stage, t_wait, t_enq);
    if (rc) {
        cudaDeviceSynchronize();
        return rc;
    }
    if (any_stats && !use_stats)
        PSTB_CUDA(cudaMemcpy(h_stats, c.d_stats.p, (size_t)n_sid * 2 * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int pstb_standardize_host(void* h_val, int dtype, int order, int64_t n_iid, int64_t n_sid, int mode, double a, double b,
                                     int apply_in_place, int use_stats, double* h_stats) {
    if (n_iid < 0 || n_sid < 0) return fail("negative shape");
    if (dtype != PSTB_F32 && dtype != PSTB_F64) return fail("standardize needs float32 or float64");
    if (n_sid == 0) return 0;
    if (!h_stats) return fail("h_stats is NULL");
    if (n_iid > 0 && !h_val) return fail("h_val is NULL");
    HostCtx& c = ctx();
    if (c.init()) return 1;
    const size_t es = esize_of(dtype);
    if (c.d_stats.ensure((size_t)n_sid * 2 * sizeof(double))) return 1;
    if (c.d_work.ensure((size_t)pstb_standardize_work_bytes(n_sid))) return 1;
    if (use_stats && upload_sync(c.d_stats.p, h_stats, (size_t)n_sid * 2 * sizeof(double), c.s[0])) return 1;
    const size_t col_bytes = (size_t)n_iid * es;
    if (order == PSTB_ORDER_C || col_bytes == 0) {
        // C order: statistics need whole columns, i.e. the whole matrix resident
        cudaStream_t st = c.s[0];
        const size_t bytes = (size_t)n_sid * col_bytes;
        if (c.d_out[0].ensure(bytes + 16)) return 1;
        if (bytes) PSTB_CUDA(cudaMemcpyAsync(c.d_out[0].p, h_val, bytes, cudaMemcpyHostToDevice, st));
        if (pstb_standardize(c.d_out[0].p, dtype, order, n_iid, n_sid, mode, a, b, apply_in_place, use_stats, (double*)c.d_stats.p,
                             c.d_work.p, st))
            return 1;
        if (apply_in_place && bytes) PSTB_CUDA(cudaMemcpyAsync(h_val, c.d_out[0].p, bytes, cudaMemcpyDeviceToHost, st));
        PSTB_CUDA(cudaStreamSynchronize(st));
    } else {
        // F order: column blocks through a kSlots-deep ring -- the H2D copy of one block, the kernel of the next and the D2H copy
        // of a third overlap (PCIe is full duplex); pageable arrays are staged through pinned buffers by host threads
        size_t target = (size_t)64 << 20;
        if (const char* e = getenv("PSTB_STD_HOST_CHUNK_KB")) { const long v = atol(e); if (v >= 1) target = (size_t)v << 10; }   // tests: many chunks
        int64_t chunk = (int64_t)(target / col_bytes);
        if (chunk < 1) chunk = 1;
        if (chunk > n_sid) chunk = n_sid;
        const bool pinned = is_pinned(h_val);
        for (int k = 0; k < kSlots; ++k) {
            if (c.d_out[k].ensure((size_t)chunk * col_bytes + 16)) return 1;
            if (!pinned && c.h_out[k].ensure((size_t)chunk * col_bytes)) return 1;
        }
        struct Pending { int64_t b0 = 0, ns = 0; bool active = false; } pend[kSlots];
        auto finish = [&](int slot) -> int {
            if (!pend[slot].active) return 0;
            PSTB_CUDA(cudaEventSynchronize(c.done[slot]));
            pend[slot].active = false;
            if (pinned || !apply_in_place) return 0;
            char* dst = (char*)h_val + (size_t)pend[slot].b0 * col_bytes;
            const char* src = (const char*)c.h_out[slot].p;
            parallel_ranges((size_t)pend[slot].ns * col_bytes, (size_t)1 << 20, [&](size_t lo, size_t hi) { memcpy(dst + lo, src + lo, hi - lo); });
            return 0;
        };
        int rc = 0;
        const int64_t nchunks = (n_sid + chunk - 1) / chunk;
        for (int64_t ch = 0; ch < nchunks && !rc; ++ch) {
            const int slot = (int)(ch % kSlots);
            const int64_t b0 = ch * chunk, ns = (b0 + chunk <= n_sid) ? chunk : n_sid - b0;
            if ((rc = finish(slot))) break;
            cudaStream_t st = c.s[slot];
            char* hp = (char*)h_val + (size_t)b0 * col_bytes;
            const size_t bytes = (size_t)ns * col_bytes;
            const void* from = hp;
            if (!pinned) {
                char* stage = (char*)c.h_out[slot].p;
                parallel_ranges(bytes, (size_t)1 << 20, [&](size_t lo, size_t hi) { memcpy(stage + lo, hp + lo, hi - lo); });
                from = stage;
            }
            if (cudaMemcpyAsync(c.d_out[slot].p, from, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = fail("H2D copy failed"); break; }
            rc = pstb_standardize(c.d_out[slot].p, dtype, order, n_iid, ns, mode, a, b, apply_in_place, use_stats,
                                  (double*)c.d_stats.p + 2 * b0, c.d_work.p, st);
            if (rc) break;
            if (apply_in_place &&
                cudaMemcpyAsync(pinned ? (void*)hp : c.h_out[slot].p, c.d_out[slot].p, bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess) {
                rc = fail("D2H copy failed");
                break;
            }
            if (cudaEventRecord(c.done[slot], st) != cudaSuccess) { rc = fail("cudaEventRecord failed"); break; }
            pend[slot].b0 = b0;
            pend[slot].ns = ns;
            pend[slot].active = true;
        }
        for (int k = 0; k < kSlots; ++k) {
            int r2 = finish(k);
            if (!rc) rc = r2;
        }
        if (rc) {
            cudaDeviceSynchronize();
            return rc;
        }
    }
    if (!use_stats) PSTB_CUDA(cudaMemcpy(h_stats, c.d_stats.p, (size_t)n_sid * 2 * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int pstb_subset_host(const void* h_in, int dtype_in, int order_in, int64_t n_in, int64_t m_in, int64_t v,
                                const int64_t* h_rows, int64_t n_rows, const int64_t* h_cols, int64_t n_cols, void* h_out,
                                int dtype_out, int order_out) {
    if (n_in < 0 || m_in < 0 || v < 0) return fail("negative shape");
    HostCtx& c = ctx();
    if (c.init()) return 1;
    cudaStream_t st = c.s[0];
    std::vector<uint32_t> scratch;
    pstb_axis rows, cols;
    Buf d_cols;
    if (make_axis(h_rows, n_rows, n_in, "row", c.d_idx, st, &rows, scratch)) return 1;
    int rc = make_axis(h_cols, n_cols, m_in, "col", d_cols, st, &cols, scratch);
    if (!rc) {
        const size_t in_bytes = (size_t)n_in * m_in * v * esize_of(dtype_in);
        const size_t out_bytes = (size_t)rows.n * cols.n * v * esize_of(dtype_out);
        if (out_bytes > 0) {
            rc = c.d_out[0].ensure(in_bytes) || c.d_out[1].ensure(out_bytes);
            if (!rc && cudaMemcpyAsync(c.d_out[0].p, h_in, in_bytes, cudaMemcpyHostToDevice, st) != cudaSuccess)
                rc = fail("H2D copy failed");
            if (!rc) rc = pstb_subset(c.d_out[0].p, dtype_in, order_in, n_in, m_in, v, rows, cols, c.d_out[1].p, dtype_out, order_out, st);
            if (!rc && cudaMemcpyAsync(h_out, c.d_out[1].p, out_bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess)
                rc = fail("D2H copy failed");
            if (cudaStreamSynchronize(st) != cudaSuccess && !rc) rc = fail("stream sync failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
    }
    d_cols.release();
    return rc;
}

// ---- SnpReader._read_kernel on host buffers (snpreader.py:623-668): packed file bytes in, K out --------------------------------
// The packed records cross PCIe in slices of a few SYRK chunks on a copy stream while the previous slice is being multiplied
// (2 bits per genotype: 6.25 GB for 50 000 x 500 000, hidden behind seconds of tensor-core work); K is accumulated on the
// device and only the finished matrix travels back, converted to the requested dtype band by band.
static int snp_kernel_host_impl(const uint8_t* h_packed, int64_t iid_count, int64_t sid_count, const int64_t* h_iid_idx,
                                int64_t n_iid, const int64_t* h_sid_idx, int64_t n_sid, int count_a1, int mode, double a, double b,
                                int use_stats, double* h_stats, void* h_K, int dtype, int64_t chunk, int low_term, bool exact) {
    if (iid_count < 0 || sid_count < 0) return fail("negative iid_count / sid_count");
    if (!h_iid_idx) n_iid = iid_count;
    if (!h_sid_idx) n_sid = sid_count;
    if (n_iid < 0 || n_sid < 0) return fail("negative selection length");
    if (iid_count > 0xfffffff0LL) return fail("iid_count too large");
    if (dtype != PSTB_F32 && dtype != PSTB_F64) return fail("kernel dtype must be float32 or float64");
    if (chunk < 64 || chunk % 64) return fail("chunk must be a positive multiple of 64");
    if (exact && dtype != PSTB_F64) return fail("the float64 kernel path returns float64");
    for (int64_t k = 0; h_sid_idx && k < n_sid; ++k)
        if (h_sid_idx[k] < 0 || h_sid_idx[k] >= sid_count)
            return fail("sid index %lld out of range [0, %lld)", (long long)h_sid_idx[k], (long long)sid_count);
    if (n_iid == 0) return 0;
    if (!h_K) return fail("h_K is NULL");
    if (n_sid > 0 && (!h_packed || !h_stats)) return fail("NULL host buffer");
    HostCtx& c = ctx();
    if (c.init()) return 1;
    std::vector<uint32_t> scratch;
    pstb_axis iid_ax;
    if (make_axis(h_iid_idx, n_iid, iid_count, "iid", c.d_idx, c.s[0], &iid_ax, scratch)) return 1;
    const bool trace = getenv("PSTB_HOST_TRACE") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (trace) fprintf(stderr, "[pstb_snp_kernel_host] %-28s %8.1f ms\n", what,
                           std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count());
    };

    const int64_t rec = (iid_count + 3) / 4, ld = pstb_packed_ld(iid_count);
    const size_t es = esize_of(dtype), kbytes32 = (size_t)n_iid * n_iid * (exact ? sizeof(double) : sizeof(float));   // device K: fp64 on the exact path
    int64_t slice = (int64_t)(((size_t)384 << 20) / (size_t)(ld > 0 ? ld : 1)) / chunk * chunk;      // ~384 MB of records per slice
    if (const char* e = getenv("PSTB_KERNEL_SLICE_SNPS")) {                                          // tests: force several slices
        const int64_t v = atoll(e);
        if (v > 0) slice = (v + chunk - 1) / chunk * chunk;
    }
    if (slice < chunk) slice = chunk;
    if (slice > n_sid) slice = (n_sid + chunk - 1) / chunk * chunk;
    const int64_t work_bytes = exact ? pstb_kernel_f64_workspace_bytes(n_iid, chunk) : pstb_kernel_workspace_bytes(n_iid, chunk);
    // the device K and the workspace stay cached in the thread's context between calls (cudaFree of a 10 GB buffer was measured
    // at up to 0.9 s); pstb_host_release() returns them
    Buf& d_K = c.d_K;
    cudaStream_t comp = c.s[0];
    cudaEvent_t* copied = c.copied;
    cudaEvent_t* used = c.used;
    int rc = 0;
    auto cleanup = [&](int r) {
        for (int k = 0; k < 3; ++k) cudaStreamSynchronize(c.s[k]);
        if (c.hp) cudaStreamSynchronize(c.hp);
        return r;
    };
    if (d_K.ensure(kbytes32) || c.d_work.ensure((size_t)work_bytes) || c.d_stats.ensure((size_t)(n_sid > 0 ? n_sid : 1) * 2 * sizeof(double)))
        return cleanup(1);
    mark("buffers allocated");
    if (use_stats && n_sid > 0 &&
        cudaMemcpyAsync(c.d_stats.p, h_stats, (size_t)n_sid * 2 * sizeof(double), cudaMemcpyHostToDevice, comp) != cudaSuccess)
        return cleanup(fail("H2D copy of the statistics failed"));
    if (n_sid == 0 && cudaMemsetAsync(d_K.p, 0, kbytes32, comp) != cudaSuccess) return cleanup(fail("cudaMemset failed"));
    const bool packed_pinned = n_sid > 0 && is_pinned(h_packed);
    bool used_pending[2] = {false, false};
    // records [b0, b0 + ns) of the selection -> d_dst (pitch ld) on copy stream c.s[1 + slot]; the copy is recorded in copied[slot]
    auto upload_records = [&](int64_t b0, int64_t ns, void* d_dst, int slot) -> int {
        cudaStream_t cp = c.s[1 + slot];
        bool contiguous = true;
        const int64_t j0 = h_sid_idx ? h_sid_idx[b0] : b0;
        for (int64_t k = 1; h_sid_idx && k < ns && contiguous; ++k) contiguous = h_sid_idx[b0 + k] == j0 + k;
        cudaError_t e = cudaSuccess;
        if (contiguous && packed_pinned) {
            if (ld == rec) {
                e = cudaMemcpyAsync(d_dst, h_packed + (size_t)j0 * rec, (size_t)ns * rec, cudaMemcpyHostToDevice, cp);
            } else {
                if (c.d_tight[slot].ensure((size_t)(ns > slice ? ns : slice) * rec)) return 1;
                e = cudaMemcpyAsync(c.d_tight[slot].p, h_packed + (size_t)j0 * rec, (size_t)ns * rec, cudaMemcpyHostToDevice, cp);
                if (e == cudaSuccess)
                    e = cudaMemcpy2DAsync(d_dst, (size_t)ld, c.d_tight[slot].p, (size_t)rec, (size_t)rec, (size_t)ns, cudaMemcpyDeviceToDevice, cp);
            }
        } else {
            // pageable or scattered records: gather them into a pinned staging buffer with host threads (the buffer is free once
            // the previous copy from it has finished)
            if (c.h_in[slot].ensure((size_t)(ns > slice ? ns : slice) * ld)) return 1;
            if (used_pending[slot]) cudaEventSynchronize(copied[slot]);
            char* stage = (char*)c.h_in[slot].p;
            parallel_ranges((size_t)ns, 64, [&](size_t lo, size_t hi) {
                for (size_t k = lo; k < hi; ++k) {
                    const int64_t j = h_sid_idx ? h_sid_idx[b0 + (int64_t)k] : b0 + (int64_t)k;
                    memcpy(stage + k * (size_t)ld, h_packed + (size_t)j * rec, (size_t)rec);
                }
            }, 4);                                                 // a slice is staged under the previous slice's SYRK (>= 10 ms): four threads are plenty
            e = cudaMemcpyAsync(d_dst, stage, (size_t)ns * ld, cudaMemcpyHostToDevice, cp);
        }
        if (e != cudaSuccess) return fail("H2D copy of packed records failed: %s", cudaGetErrorString(e));
        if (cudaEventRecord(copied[slot], cp) != cudaSuccess) return fail("cudaEventRecord failed");
        used_pending[slot] = true;                                  // (the staging buffer / d_tight of this slot are in use until copied[slot])
        return 0;
    };

    // ---- K back to the host: row bands, converted on the device, D2H overlapped with the next band's conversion ----
    // pinned destination: 2 slots of 128 MiB straight into it; pageable one: up to 16 slots of 32 MiB, each drained by one worker of
    // the calling thread's task pool (as in pstb_read_host)
    const bool out_pinned = is_pinned(h_K);
    const size_t row_bytes = (size_t)n_iid * es;
    const int nsl = out_pinned ? 2 : kMaxSlots;
    int64_t band = (int64_t)(((size_t)(out_pinned ? 128 : 32) << 20) / row_bytes);
    if (band < 1) band = 1;
    if (band > n_iid) band = n_iid;
    struct Pend { bool active = false; } pend[kMaxSlots];
    std::mutex slot_m;
    std::condition_variable slot_cv;
    bool slot_busy[kMaxSlots] = {};
    std::atomic<int> worker_rc{0};
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    if (!out_pinned) task_pool().ensure(host_copy_threads());
    int64_t drain_seq = 0;
    auto finish = [&](int slot) -> int {
        if (!pend[slot].active) return 0;
        pend[slot].active = false;
        if (!out_pinned) {
            std::unique_lock<std::mutex> lk(slot_m);
            slot_cv.wait(lk, [&] { return !slot_busy[slot]; });
            return worker_rc.load() ? fail("a host copy worker failed (CUDA error %d)", worker_rc.load()) : 0;
        }
        if (cudaEventSynchronize(c.done[slot]) != cudaSuccess) return fail("D2H copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        return 0;
    };
    // rows [ra, rb) of the finished device K -> the host.  `after` (may be NULL): the event that makes these rows final; `side`:
    // copy on streams 1 / 2 only (the compute stream is still busy with other rows)
    auto drain_rows = [&](int64_t ra, int64_t rb, cudaEvent_t after, bool side) -> int {
        for (int64_t r0 = ra; r0 < rb; r0 += band, ++drain_seq) {
            const int slot = (int)(drain_seq % nsl);
            const int64_t nr = (r0 + band <= rb) ? band : rb - r0;
            if (int r = finish(slot)) return r;
            cudaStream_t st = side ? c.s[1 + slot % 2] : c.s[slot % 2];
            if (after && cudaStreamWaitEvent(st, after, 0) != cudaSuccess) return fail("cudaStreamWaitEvent failed");
            const float* src = (const float*)d_K.p + (size_t)r0 * n_iid;
            const void* from = src;
            if (exact) {
                from = (const double*)d_K.p + (size_t)r0 * n_iid;           // already float64: straight D2H
            } else if (dtype == PSTB_F64) {
                if (c.d_out[slot].ensure((size_t)band * row_bytes)) return 1;
                if (int r = convert_range(src, (long long)nr * n_iid, c.d_out[slot].p, dtype, 1.0, st)) return r;
                from = c.d_out[slot].p;
            }
            void* dst = (char*)h_K + (size_t)r0 * row_bytes;
            if (!out_pinned) {
                if (c.h_out[slot].ensure((size_t)band * row_bytes)) return 1;
                dst = c.h_out[slot].p;
            }
            if (cudaMemcpyAsync(dst, from, (size_t)nr * row_bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                cudaEventRecord(c.done[slot], st) != cudaSuccess)
                return fail("D2H copy of K failed");
            pend[slot].active = true;
            if (!out_pinned) {
                {
                    std::unique_lock<std::mutex> lk(slot_m);
                    slot_busy[slot] = true;
                }
                cudaEvent_t ev = c.done[slot];
                const char* stage = (const char*)c.h_out[slot].p;
                char* dest = (char*)h_K + (size_t)r0 * row_bytes;
                const size_t bytes = (size_t)nr * row_bytes;
                task_pool().submit([&, slot, ev, stage, dest, bytes, cur_dev] {
                    static thread_local int dev_set = -1;
                    if (dev_set != cur_dev) { cudaSetDevice(cur_dev); dev_set = cur_dev; }
                    const cudaError_t e = cudaEventSynchronize(ev);
                    if (e != cudaSuccess) worker_rc.store((int)e); else memcpy(dest, stage, bytes);
                    {
                        std::unique_lock<std::mutex> lk(slot_m);
                        slot_busy[slot] = false;
                    }
                    slot_cv.notify_all();
                });
            }
        }
        return 0;
    };
    auto finish_all = [&](int r) -> int {                           // no worker may outlive this frame, whatever happened
        for (int k = 0; k < kMaxSlots; ++k) {
            const int r2 = finish(k);
            if (!r) r = r2;
        }
        return r;
    };

    // ---- K final only after the last SNP chunk: its 10 GB (cfg3) left over PCIe AFTER the multiplication, 0.18 s of a 1.9 s call.
    // Overlapped copy-out: K accumulates in compact lower-triangular tiles; the last T chunks (as many as the copy-out takes) are
    // multiplied BAND-major, from the bottom band of tiles up: once a band's tiles are final and expanded (both orientations + the
    // rank-one part), the rows of the square K below its first tile row are complete and leave while the bands above still multiply.
    const int64_t nchunks = (n_sid + chunk - 1) / chunk;
    const int64_t ntiles = pstb_kernel_tile_count(n_iid, 0, 1);
    int overlap_mode = 1;
    if (const char* e = getenv("PSTB_HOST_KERNEL_OVERLAP")) overlap_mode = atoi(e);     // 0: off, 1: large kernels, 2: whenever possible (tests)
    bool overlap = !exact && overlap_mode > 0 && nchunks >= 2 && ntiles >= 3 && n_iid >= 512 &&
                   (overlap_mode >= 2 || (n_iid >= 8192 && nchunks >= 8));
    int64_t tail_chunks = 0;
    if (overlap) {
        // tail length: the copy-out moves n^2 * es bytes at ~50 GB/s (a fresh pageable destination takes them at ~35: first-touch page
        // faults), a chunk multiplies 2 n^2 chunk flop at ~1.45e15/s
        // (a tail 1.7 x as long measured 10 ms SLOWER on cfg3: the band-major launches are shorter and less efficient than whole chunks)
        double tail_scale = 1.0;
        if (const char* e = getenv("PSTB_HOST_KERNEL_TAIL")) { const double v = atof(e); if (v >= 0.1 && v <= 8.0) tail_scale = v; }   // tuning experiments
        int64_t T = (int64_t)(tail_scale * (double)es * (out_pinned ? 14500.0 : 21000.0) / (double)chunk) + 1;
        if (T > 32) T = 32;
        while (T > 1 && (size_t)T * (size_t)work_bytes > ((size_t)40 << 30)) --T;
        if (T > nchunks) T = nchunks;
        // the extra device memory (compact tiles + T workspaces + the tail's records) must fit beside K: shorten the tail, or take the
        // plain loop, rather than fail
        {
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = 0; }
            const size_t have = c.d_tiles.cap + c.d_tail_work.cap + c.d_tail_packed.cap;          // cached from an earlier call: reused or released by ensure()
            const size_t budget = free_b + have > ((size_t)3 << 30) ? free_b + have - ((size_t)3 << 30) : 0;
            const size_t tiles_b = (size_t)ntiles * 65536 * sizeof(float);
            while (T >= 1 && tiles_b + (size_t)T * ((size_t)work_bytes + (size_t)chunk * (size_t)ld) > budget) --T;
            if (T < 1) overlap = false;
        }
        tail_chunks = T;
    }
    if (overlap) {
        int64_t T = tail_chunks;
        const int64_t head = (nchunks - T) * chunk;                // SNPs multiplied chunk-major, slice by slice
        const int low = pstb_resolve_low_term(low_term, n_sid, n_iid, mode);
        const size_t n_pad2 = ((size_t)n_iid + 255) / 256 * 256 + 2;
        if (c.d_tiles.ensure((size_t)ntiles * 65536 * sizeof(float)) || c.d_tail_work.ensure((size_t)T * (size_t)work_bytes) ||
            c.d_tail_packed.ensure((size_t)T * (size_t)chunk * (size_t)ld) || c.d_u.ensure(n_pad2 * sizeof(double)))
            return cleanup(1);
        // bands: cuts of the tile list where every later tile lies in a lower tile row than every earlier one
        std::vector<int32_t> ij((size_t)ntiles * 2);
        if (pstb_kernel_tile_coords(n_iid, 0, 1, ij.data())) return cleanup(1);
        std::vector<int32_t> sufmin((size_t)ntiles + 1, INT32_MAX);
        for (int64_t t = ntiles - 1; t >= 0; --t) sufmin[(size_t)t] = ij[(size_t)2 * t] < sufmin[(size_t)t + 1] ? ij[(size_t)2 * t] : sufmin[(size_t)t + 1];
        std::vector<int64_t> cuts;                                   // band starts, ascending; cuts[0] = 0
        cuts.push_back(0);
        {
            // finer than the reduction's bands (the LAST band's rows leave exposed): every super-block row of tiles at cfg3's size.
            // Measured e2e minus device-resident time, cfg3: 12 bands 85 ms, 100 (= all 25 possible cuts) 72 ms
            int64_t parts = 64;
            if (const char* e = getenv("PSTB_HOST_KERNEL_BANDS")) { const int v = atoi(e); if (v >= 1 && v <= 1024) parts = v; }   // tuning experiments
            const int64_t want = ntiles / parts > 1 ? ntiles / parts : 1;
            int32_t premax = -1;
            for (int64_t t = 1; t < ntiles; ++t) {
                premax = ij[(size_t)2 * (t - 1)] > premax ? ij[(size_t)2 * (t - 1)] : premax;
                if (premax < sufmin[(size_t)t] && t - cuts.back() >= want) cuts.push_back(t);
            }
        }
        cuts.push_back(ntiles);
        const int nbands = (int)cuts.size() - 1;
        std::vector<cudaEvent_t> ev_band((size_t)nbands, nullptr), ev_exp((size_t)nbands, nullptr);
        auto destroy_events = [&]() {
            for (auto& e : ev_band) if (e) cudaEventDestroy(e);
            for (auto& e : ev_exp) if (e) cudaEventDestroy(e);
        };
        double* d_u = (double*)c.d_u.p;
        if (cudaMemsetAsync(d_u, 0, n_pad2 * sizeof(double), comp) != cudaSuccess) return cleanup(fail("cudaMemset failed"));
        auto add_rank1 = [&](void* work) -> int {
            const double* u = pstb_kernel_workspace_rank1(work, n_iid, chunk);
            if (!u) return fail("no rank-one vector in the workspace");
            k_add_f64<<<(unsigned)((n_iid + 255) / 256), 256, 0, comp>>>(d_u, u, (long long)n_iid);
            return cudaGetLastError() == cudaSuccess ? 0 : fail("k_add_f64 launch failed");
        };
        // the tail's records first (one short copy), then the head slice by slice as in the plain loop
        const int64_t tail_sid = n_sid - head;
        for (int64_t b0 = head, piece = 0; b0 < n_sid && !rc; b0 += slice, ++piece) {
            const int64_t ns = (b0 + slice <= n_sid) ? slice : n_sid - b0;
            rc = upload_records(b0, ns, (char*)c.d_tail_packed.p + (size_t)(b0 - head) * ld, (int)(piece & 1));
        }
        if (rc) return cleanup(rc);
        for (int64_t b0 = 0, sl = 0; b0 < head && !rc; b0 += slice, ++sl) {
            const int slot = (int)(sl & 1);
            const int64_t ns = (b0 + slice <= head) ? slice : head - b0;
            if (c.d_packed[slot].ensure((size_t)slice * ld)) { rc = 1; break; }
            if (used_pending[slot] && cudaStreamWaitEvent(c.s[1 + slot], used[slot], 0) != cudaSuccess) { rc = fail("cudaStreamWaitEvent failed"); break; }
            if ((rc = upload_records(b0, ns, c.d_packed[slot].p, slot))) break;
            if (cudaStreamWaitEvent(comp, copied[slot], 0) != cudaSuccess) { rc = fail("cudaStreamWaitEvent failed"); break; }
            pstb_axis sid_ax{nullptr, 0, 1, ns};
            rc = pstb_snp_kernel_tiles((const uint8_t*)c.d_packed[slot].p, ld, iid_count, ns, iid_ax, sid_ax, count_a1, mode, a, b, use_stats,
                                       (double*)c.d_stats.p + 2 * b0, (float*)c.d_tiles.p, 0, 1, (b0 > 0 ? 1 : 0) | 2, c.d_work.p, work_bytes, chunk, low, comp);
            if (rc || (rc = add_rank1(c.d_work.p))) break;
            if (cudaEventRecord(used[slot], comp) != cudaSuccess) { rc = fail("cudaEventRecord failed"); break; }
        }
        if (rc) return cleanup(rc);
        mark("head enqueued");
        // the tail's records went first on the two copy streams: the latest event of each stream covers them (the head slices above
        // did not have to wait for that copy)
        for (int k = 0; k < 2; ++k)
            if (used_pending[k] && cudaStreamWaitEvent(comp, copied[k], 0) != cudaSuccess) return cleanup(fail("cudaStreamWaitEvent failed"));
        auto band_call = [&](int64_t tc, int64_t t0, int64_t t1, int flags) -> int {
            const int64_t lo = tc * chunk, ns = (lo + chunk <= tail_sid) ? chunk : tail_sid - lo;
            pstb_axis sid_ax{nullptr, 0, 1, ns};
            return pstb_snp_kernel_tiles_band((const uint8_t*)c.d_tail_packed.p + (size_t)lo * ld, ld, iid_count, ns, iid_ax, sid_ax, count_a1, mode, a, b,
                                              use_stats, (double*)c.d_stats.p + 2 * (head + lo), (float*)c.d_tiles.p, 0, 1, (head > 0 || tc > 0) ? 1 : 0,
                                              (char*)c.d_tail_work.p + (size_t)tc * (size_t)work_bytes, work_bytes, chunk, low, t0, t1, flags, 0, comp);
        };
        for (int64_t tc = 0; tc < T && !rc; ++tc) {                 // statistics + operand planes of the tail chunks, no tiles yet
            if ((rc = band_call(tc, 0, 0, 1 | 2))) break;
            rc = add_rank1((char*)c.d_tail_work.p + (size_t)tc * (size_t)work_bytes);
        }
        // every band's multiplication and expansion is enqueued first (the copy-out below blocks the caller on its staging slots)
        for (int bnd = nbands - 1; bnd >= 0 && !rc; --bnd) {
            const int64_t t0 = cuts[(size_t)bnd], t1 = cuts[(size_t)bnd + 1];
            for (int64_t tc = 0; tc < T && !rc; ++tc) rc = band_call(tc, t0, t1, 2);
            if (rc) break;
            if (cudaEventCreateWithFlags(&ev_band[(size_t)bnd], cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&ev_exp[(size_t)bnd], cudaEventDisableTiming) != cudaSuccess ||
                cudaEventRecord(ev_band[(size_t)bnd], comp) != cudaSuccess || cudaStreamWaitEvent(c.hp, ev_band[(size_t)bnd], 0) != cudaSuccess) {
                rc = fail("band event failed");
                break;
            }
            if ((rc = pstb_kernel_from_tiles_range((const float*)c.d_tiles.p, n_iid, 0, 1, t0, t1, (float*)d_K.p, d_u, c.hp))) break;
            if (cudaEventRecord(ev_exp[(size_t)bnd], c.hp) != cudaSuccess) rc = fail("cudaEventRecord failed");
        }
        mark("tail enqueued");
        int64_t row_hi = n_iid;
        for (int bnd = nbands - 1; bnd >= 0 && !rc; --bnd) {
            int64_t row_lo = bnd == 0 ? 0 : (int64_t)sufmin[(size_t)cuts[(size_t)bnd]] * 256;
            if (row_lo > n_iid) row_lo = n_iid;
            if (row_lo < row_hi) {
                rc = drain_rows(row_lo, row_hi, ev_exp[(size_t)bnd], true);
                row_hi = row_lo;
            }
        }
        rc = finish_all(rc);
        if (cudaStreamSynchronize(comp) != cudaSuccess || cudaStreamSynchronize(c.hp) != cudaSuccess)
            rc = rc ? rc : fail("kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
        destroy_events();
        mark("K on the host");
        if (!rc && !use_stats && n_sid > 0 &&
            cudaMemcpy(h_stats, c.d_stats.p, (size_t)n_sid * 2 * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
            rc = fail("D2H copy of the statistics failed");
        return cleanup(rc);
    }

    for (int64_t b0 = 0, sl = 0; b0 < n_sid && !rc; b0 += slice, ++sl) {
        const int slot = (int)(sl & 1);
        const int64_t ns = (b0 + slice <= n_sid) ? slice : n_sid - b0;
        if (c.d_packed[slot].ensure((size_t)slice * ld)) { rc = 1; break; }
        if (used_pending[slot] && cudaStreamWaitEvent(c.s[1 + slot], used[slot], 0) != cudaSuccess) { rc = fail("cudaStreamWaitEvent failed"); break; }
        if ((rc = upload_records(b0, ns, c.d_packed[slot].p, slot))) break;
        if (cudaStreamWaitEvent(comp, copied[slot], 0) != cudaSuccess) { rc = fail("event record / wait failed"); break; }
        pstb_axis sid_ax{nullptr, 0, 1, ns};
        if (exact)
            rc = snp_kernel_f64_slice((const uint8_t*)c.d_packed[slot].p, ld, iid_count, ns, iid_ax, sid_ax, count_a1, mode, a, b, use_stats,
                                      (double*)c.d_stats.p + 2 * b0, (double*)d_K.p, b0 > 0 ? 1 : 0, c.d_work.p, work_bytes, chunk, comp);
        else
            rc = snp_kernel_slice((const uint8_t*)c.d_packed[slot].p, ld, iid_count, ns, iid_ax, sid_ax, count_a1, mode, a, b, use_stats,
                                  (double*)c.d_stats.p + 2 * b0, (float*)d_K.p, b0 > 0 ? 1 : 0, c.d_work.p, work_bytes, chunk, low_term, comp,
                                  (b0 == 0 ? 1 : 0) | (b0 + slice >= n_sid ? 2 : 0), n_sid);
        if (rc) break;
        if (cudaEventRecord(used[slot], comp) != cudaSuccess) { rc = fail("cudaEventRecord failed"); break; }
    }
    if (rc) return cleanup(rc);
    mark("slices enqueued");
    if (exact ? pstb_mirror_lower_f64((double*)d_K.p, n_iid, n_iid, comp) : pstb_mirror_lower((float*)d_K.p, n_iid, n_iid, comp)) return cleanup(1);
    if (cudaStreamSynchronize(comp) != cudaSuccess) return cleanup(fail("kernel failed: %s", cudaGetErrorString(cudaGetLastError())));
    mark("kernel finished");
    rc = drain_rows(0, n_iid, nullptr, false);
    rc = finish_all(rc);
    mark("K on the host");
    if (!rc && !use_stats && n_sid > 0 &&
        cudaMemcpy(h_stats, c.d_stats.p, (size_t)n_sid * 2 * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
        rc = fail("D2H copy of the statistics failed");
    rc = cleanup(rc);
    mark("buffers released");
    return rc;
}

extern "C" int pstb_snp_kernel_host(const uint8_t* h_packed, int64_t iid_count, int64_t sid_count, const int64_t* h_iid_idx,
                                    int64_t n_iid, const int64_t* h_sid_idx, int64_t n_sid, int count_a1, int mode, double a, double b,
                                    int use_stats, double* h_stats, void* h_K, int dtype, int64_t chunk, int low_term) {
    return snp_kernel_host_impl(h_packed, iid_count, sid_count, h_iid_idx, n_iid, h_sid_idx, n_sid, count_a1, mode, a, b, use_stats, h_stats, h_K,
                                dtype, chunk, low_term, false);
}

// the same loop in float64 arithmetic (syrk_f64.cu): what a dtype=float64 request of the reference means (snpdata.py:203-206 is a DGEMM then)
extern "C" int pstb_snp_kernel_host_f64(const uint8_t* h_packed, int64_t iid_count, int64_t sid_count, const int64_t* h_iid_idx,
                                        int64_t n_iid, const int64_t* h_sid_idx, int64_t n_sid, int count_a1, int mode, double a, double b,
                                        int use_stats, double* h_stats, double* h_K, int64_t chunk) {
    return snp_kernel_host_impl(h_packed, iid_count, sid_count, h_iid_idx, n_iid, h_sid_idx, n_sid, count_a1, mode, a, b, use_stats, h_stats, h_K,
                                PSTB_F64, chunk, PSTB_LOW_TERM_DEFAULT, true);
}

// free every device / pinned buffer the host-buffer entry points keep cached for the calling thread
extern "C" int pstb_host_release(void) {
    HostCtx& c = ctx();
    if (c.device < 0) return 0;
    cudaDeviceSynchronize();
    c.release_buffers();
    return 0;
}
