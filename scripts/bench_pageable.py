"""pstb_read_host into a FRESH pageable NumPy array (what Bed.read returns) for several host copy-thread counts (experiment)."""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pysnptools_b200 import _lib
lib = _lib.lib
n, m = 10000, int(sys.argv[1]) if len(sys.argv) > 1 else 250000
rec = (n + 3) // 4
pk = lib.pstb_host_alloc(m * rec)
np.ctypeslib.as_array(ctypes.cast(pk, ctypes.POINTER(ctypes.c_uint8)), shape=(m * rec,))[:] = 0x9c
st = np.empty((m, 2))
def step(out):
    _lib.check(lib.pstb_read_host(pk, n, m, None, n, None, m, 0, 1, float("nan"), float("nan"), 0, st.ctypes.data, out.ctypes.data, 0, 0))
warm = np.empty((n, 4096), dtype=np.float32, order="F")
_lib.check(lib.pstb_read_host(pk, n, 4096, None, n, None, 4096, 0, 1, float("nan"), float("nan"), 0, st.ctypes.data, warm.ctypes.data, 0, 0))
for threads in [None, 4, 8, 12, 16, 24, 32]:
    if threads is None:
        os.environ.pop("PSTB_HOST_COPY_THREADS", None)
    else:
        os.environ["PSTB_HOST_COPY_THREADS"] = str(threads)
    ts = []
    for rep in range(2):
        t0 = time.perf_counter()
        out = np.empty((n, m), dtype=np.float32, order="F")
        step(out)
        ts.append(time.perf_counter() - t0)
        t1 = time.perf_counter()
        step(out)                                      # second pass: pages already faulted in
        warm_t = time.perf_counter() - t1
        del out
    print("copy threads %-8s fresh array: %.3f s = %.2e genotypes/s (%.1f GB/s)   pre-faulted array: %.3f s (%.1f GB/s)" % (
        threads or "default", min(ts), n * m / min(ts), n * m * 4 / min(ts) / 1e9, warm_t, n * m * 4 / warm_t / 1e9), flush=True)
