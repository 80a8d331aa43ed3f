"""Where the time of Bed(file).read(dtype=float32, standardizer=Unit(), out=pinned) goes beyond pstb_read_host on pinned buffers (cfg2 shape):
the same C call with the packed bytes (a) in page-locked memory, (b) in an ordinary NumPy array, (c) in the memory-mapped file,
with PSTB_HOST_TRACE=1 (the library prints its wait / enqueue split), then the Python call under cProfile."""
import cProfile, ctypes, os, pstats, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
os.environ["PSTB_HOST_TRACE"] = "1"
from pysnptools_b200 import _lib, Bed, Unit
from pysnptools_b200.util import pinned_empty
lib = _lib.lib
n, m = 10000, int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
rec = (n + 3) // 4
rng = np.random.default_rng(0)
packed = rng.integers(0, 256, size=(m, rec), dtype=np.uint8)
packed &= ~((packed & 0x55) & ~((packed >> 1) & 0x55))          # no missing codes
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


pin = pinned_empty((m, rec), dtype=np.uint8, order="C")
pin[...] = packed
call("packed bytes page-locked", pin)
call("packed bytes in a NumPy array", packed)
d = tempfile.mkdtemp(prefix="pstb_prof_")
path = os.path.join(d, "x.bed")
with open(path, "wb") as f:
    f.write(bytes([0x6C, 0x1B, 0x01])); f.write(packed.tobytes())
mm = np.memmap(path, dtype=np.uint8, mode="r", offset=3, shape=(m, rec))
call("packed bytes in the memory-mapped file", mm)
bed = Bed(path, count_A1=False, iid=np.array([["f", str(k)] for k in range(n)]), sid=np.arange(m).astype(str), pos=np.zeros((m, 3)))
for _ in range(2):
    t0 = time.perf_counter(); bed.read(order="F", dtype=np.float32, standardizer=Unit(), out=out); print("Bed.read(out=pinned): %.3f s" % (time.perf_counter() - t0), flush=True)
for _ in range(3):
    t0 = time.perf_counter(); dd = bed.read(order="F", dtype=np.float32, standardizer=Unit()); dt = time.perf_counter() - t0
    print("Bed.read() -> fresh pageable array: %.3f s = %.3e genotypes/s" % (dt, n * m / dt), flush=True); del dd
pr = cProfile.Profile(); pr.enable(); bed.read(order="F", dtype=np.float32, standardizer=Unit(), out=out); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
os.remove(path); os.rmdir(d)
