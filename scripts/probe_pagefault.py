"""Host-side probe for the pageable-destination path of pstb_read_host (not product code): how fast can a FRESH NumPy array be filled
on this box?  Prints the transparent-huge-page settings and the fill rate (GB/s) of a fresh 8 GiB array for several thread counts,
plain / after madvise(MADV_HUGEPAGE) / after MADV_POPULATE_WRITE, and the cost of cudaHostRegister."""
import ctypes
import os
import sys
import threading
import time

import numpy as np

for f in ("enabled", "defrag", "shmem_enabled"):
    try:
        print("THP", f, open("/sys/kernel/mm/transparent_hugepage/" + f).read().strip())
    except Exception as e:
        print("THP", f, "unreadable", e)
print("cpus", os.cpu_count(), "numpy", np.__version__, "NUMPY_MADVISE_HUGEPAGE", os.environ.get("NUMPY_MADVISE_HUGEPAGE"))
try:
    print(open("/proc/meminfo").read().split("\n")[0:3])
except Exception:
    pass
libc = ctypes.CDLL("libc.so.6", use_errno=True)
libc.madvise.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
MADV_HUGEPAGE, MADV_NOHUGEPAGE, MADV_POPULATE_WRITE = 14, 15, 23
SIZE = int(sys.argv[1]) << 30 if len(sys.argv) > 1 else 8 << 30
src = np.full(256 << 20, 7, dtype=np.uint8)


def fill(dst, threads):
    n = dst.size
    per = (n + threads - 1) // threads

    def work(t):
        lo, hi = t * per, min(n, (t + 1) * per)
        for o in range(lo, hi, src.size):
            e = min(hi, o + src.size)
            dst[o:e] = src[: e - o]
    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    [x.start() for x in th]
    [x.join() for x in th]
    return time.perf_counter() - t0


def madvise(arr, advice, threads=1):
    a = arr.ctypes.data
    lo = (a + 4095) & ~4095
    hi = (a + arr.nbytes) & ~4095
    if threads == 1:
        r = libc.madvise(lo, hi - lo, advice)
        return r, ctypes.get_errno()
    per = ((hi - lo) // threads + (2 << 20) - 1) & ~((2 << 20) - 1)
    res = []

    def work(t):
        s = lo + t * per
        e = min(hi, s + per)
        if s < e:
            res.append(libc.madvise(s, e - s, advice))
    th = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    [x.start() for x in th]
    [x.join() for x in th]
    return max(res) if res else 0, ctypes.get_errno()


for threads in (4, 8, 16, 32):
    d = np.empty(SIZE, dtype=np.uint8)
    dt = fill(d, threads)
    dt2 = fill(d, threads)
    print("fresh np.empty fill, %2d threads: %.2f s = %5.1f GB/s   (second pass over the same pages: %5.1f GB/s)" % (threads, dt, SIZE / dt / 1e9, SIZE / dt2 / 1e9), flush=True)
    del d
for threads in (8, 16):
    d = np.empty(SIZE, dtype=np.uint8)
    r = madvise(d, MADV_NOHUGEPAGE)
    dt = fill(d, threads)
    print("MADV_NOHUGEPAGE (rc %s) then fill, %2d threads: %5.1f GB/s" % (r, threads, SIZE / dt / 1e9), flush=True)
    del d
    d = np.empty(SIZE, dtype=np.uint8)
    r = madvise(d, MADV_HUGEPAGE)
    dt = fill(d, threads)
    print("MADV_HUGEPAGE   (rc %s) then fill, %2d threads: %5.1f GB/s" % (r, threads, SIZE / dt / 1e9), flush=True)
    del d
    d = np.empty(SIZE, dtype=np.uint8)
    madvise(d, MADV_HUGEPAGE)
    t0 = time.perf_counter()
    r = madvise(d, MADV_POPULATE_WRITE, threads)
    tp = time.perf_counter() - t0
    dt = fill(d, threads)
    print("MADV_POPULATE_WRITE with %2d threads (rc %s): %.2f s = %5.1f GB/s, then fill %5.1f GB/s" % (threads, r, tp, SIZE / tp / 1e9, SIZE / dt / 1e9), flush=True)
    del d
try:
    import torch
    rt = torch.cuda.cudart()
    d = np.empty(SIZE, dtype=np.uint8)
    t0 = time.perf_counter()
    rc = rt.cudaHostRegister(d.ctypes.data, d.nbytes, 0)
    t1 = time.perf_counter()
    print("cudaHostRegister of a fresh array: rc %s, %.2f s = %.1f GB/s" % (rc, t1 - t0, SIZE / (t1 - t0) / 1e9), flush=True)
    g = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    h = torch.from_numpy(d)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(SIZE >> 30):
        h[k << 30:(k + 1) << 30].copy_(g, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("D2H into the registered array: %.1f GB/s" % (SIZE / dt / 1e9))
    rt.cudaHostUnregister(d.ctypes.data)
except Exception as e:
    print("cudaHostRegister probe failed:", e)

# ---- the same fills next to a saturated D2H DMA stream (what pstb_read_host's staging copy competes with) ----
try:
    import torch
    g = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    hp = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
    stop = False
    moved = [0]

    def dma():
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            while not stop:
                hp.copy_(g, non_blocking=True)
                s.synchronize()
                moved[0] += 1
    th = threading.Thread(target=dma)
    th.start()
    time.sleep(0.2)
    for threads in (8, 14, 16):
        d = np.empty(SIZE, dtype=np.uint8)
        m0, t0 = moved[0], time.perf_counter()
        dt = fill(d, threads)
        dma_rate = (moved[0] - m0) * (1 << 30) / (time.perf_counter() - t0) / 1e9
        dt2 = fill(d, threads)
        print("next to a D2H DMA stream (%.0f GB/s): fresh fill %2d threads %5.1f GB/s, second pass %5.1f GB/s" % (dma_rate, threads, SIZE / dt / 1e9, SIZE / dt2 / 1e9), flush=True)
        del d
    # copying OUT OF the pinned buffer the DMA writes (cold lines, as in the pipeline) instead of a resident source
    src_pinned = hp.numpy()
    for threads in (8, 14):
        d = np.empty(SIZE, dtype=np.uint8)
        src_backup = src
        globals()["src"] = src_pinned[: 256 << 20]
        dt = fill(d, threads)
        dt2 = fill(d, threads)
        globals()["src"] = src_backup
        print("source = the pinned DMA target: fresh fill %2d threads %5.1f GB/s, second pass %5.1f GB/s" % (threads, SIZE / dt / 1e9, SIZE / dt2 / 1e9), flush=True)
        del d
    stop = True
    th.join()
except Exception as e:
    print("DMA contention probe failed:", e)
