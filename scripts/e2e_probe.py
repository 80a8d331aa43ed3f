"""PCIe ceiling vs pstb_read_host (experiment)."""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pysnptools_b200 import _lib
lib = _lib.lib
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
for _ in range(2): h.copy_(x, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(8): h.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("pinned D2H 8 x 1 GiB: %.1f GB/s" % (8 * (1 << 30) / dt / 1e9))
t0 = time.perf_counter()
for _ in range(8): x.copy_(h, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("pinned H2D 8 x 1 GiB: %.1f GB/s" % (8 * (1 << 30) / dt / 1e9))
n, m = 10000, 250000
rec = 2500
pk = lib.pstb_host_alloc(m * rec); out = lib.pstb_host_alloc(n * m * 4)
np.ctypeslib.as_array(ctypes.cast(pk, ctypes.POINTER(ctypes.c_uint8)), shape=(m * rec,))[:] = 0x9c
st = np.empty((m, 2))
def step(): _lib.check(lib.pstb_read_host(pk, n, m, None, n, None, m, 0, 1, float("nan"), float("nan"), 0, st.ctypes.data, out, 0, 0))
step()
t0 = time.perf_counter(); step(); step(); dt = (time.perf_counter() - t0) / 2
print("pstb_read_host %d x %d (chunk %s MB): %.3f s  %.1f GB/s D2H  %.3e genotypes/s" % (n, m, os.environ.get("PSTB_HOST_CHUNK_MB", "64"), dt, n * m * 4 / dt / 1e9, n * m / dt))
