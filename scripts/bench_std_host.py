"""standardize_f32 through the host-buffer C ABI (what the bed_reader shim calls): pageable and pinned arrays."""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pysnptools_b200 import _lib
from pysnptools_b200.util import pinned_empty
lib = _lib.lib
n, m = 10000, 250000                                            # 10 GB float32
rng = np.random.default_rng(0)
col = rng.integers(0, 3, size=n).astype(np.float32)
for name, val in (("pageable", np.empty((n, m), dtype=np.float32, order="F")), ("pinned", pinned_empty((n, m), dtype=np.float32, order="F"))):
    val[...] = col[:, None]
    st = np.zeros((m, 2))
    for rep in range(3):
        t0 = time.perf_counter()
        rc = lib.pstb_standardize_host(val.ctypes.data, _lib.F32, 0, n, m, 1, float("nan"), float("nan"), 1, 0, st.ctypes.data)
        dt = time.perf_counter() - t0
        assert rc == 0, _lib.last_error()
        print("standardize_f32 host %s 10000 x 250000: %.3f s  %.1f GB/s each way  %.3e values/s" % (name, dt, 4.0 * n * m / dt / 1e9, n * m / dt), flush=True)
        val[...] = col[:, None]
    del val
