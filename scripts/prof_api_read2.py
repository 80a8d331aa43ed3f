"""pstb_read_host, pageable packed input + pinned output (cfg2 shape): copy-thread count sweep (PSTB_HOST_COPY_THREADS is read per call)."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
os.environ["PSTB_HOST_TRACE"] = "1"
from pysnptools_b200 import _lib
from pysnptools_b200.util import pinned_empty
lib = _lib.lib
n, m = 10000, 1000000
rec = (n + 3) // 4
packed = np.random.default_rng(0).integers(0, 256, size=(m, rec), dtype=np.uint8)
packed &= ~((packed & 0x55) & ~((packed >> 1) & 0x55))
out = pinned_empty((n, m), dtype=np.float32, order="F")
stats = np.empty((m, 2))
p = ctypes.c_void_p
def call(tag, src):
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        rc = lib.pstb_read_host(p(src.ctypes.data), n, m, None, n, None, m, 0, _lib.STD_UNIT, float("nan"), float("nan"), 0, p(stats.ctypes.data),
                                p(out.ctypes.data), _lib.F32, _lib.ORDER_F)
        ts.append(time.perf_counter() - t0)
        assert rc == 0, _lib.last_error()
    print("%-40s %s  -> %.3e genotypes/s" % (tag, ["%.3f" % t for t in ts], n * m / min(ts)), flush=True)
for th in sys.argv[1:] or ["1", "2", "4", "8", "14"]:
    os.environ["PSTB_HOST_COPY_THREADS"] = th
    call("pageable in, pinned out, %s copy threads" % th, packed)
del os.environ["PSTB_HOST_COPY_THREADS"]
pin = pinned_empty((m, rec), dtype=np.uint8, order="C"); pin[...] = packed
call("pinned in, pinned out", pin)
