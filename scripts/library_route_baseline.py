"""The reference's OWN GPU route restated with library calls, timed on the same B200 (second baseline).

ARRAY_MODULE=cupy in the reference: host decode (Rust) -> H2D of the float matrix (unit.py:38) -> generic element-wise /
reduction kernels for the standardize (standardizer.py:145-163) -> cupy.dot = cuBLAS in the read dtype (snpdata.py:203-206)
-> K += on the device (snpreader.py:655).  Here with torch ops (CuPy is not installed): nanmean / nanstd-style standardize +
torch.matmul in float64 (the reference default dtype), float32 and TF32.  Not product code; numbers go into DESIGN.md.
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pysnptools_b200 import device as dev

def standardize_lib(x):                       # the six-sweep library restatement of _standardize_unit_python
    miss = torch.isnan(x)
    mean = torch.nanmean(x, dim=0)
    dev_ = torch.where(miss, torch.zeros_like(x), x - mean)
    n = (~miss).sum(dim=0)
    std = torch.sqrt((dev_ * dev_).sum(dim=0) / n)
    std[std == 0] = float("inf")
    x -= mean
    x /= std
    x[miss] = 0
    return x

def ev_time(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

n, block = 50000, 10000                        # cfg3 height, the reference's block_size-style SNP blocks
store = bench.gen_store_device(dev, torch, n, block, seed=1)
raw32, _ = dev.read(store, dtype=np.float32, order="C")           # decode is done for the library route (it happens on the CPU there)
for name, dtype, tf32 in (("float64 (reference default)", torch.float64, False), ("float32", torch.float32, False), ("float32 with TF32 tensor cores", torch.float32, True)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    x = raw32.to(dtype)
    K = torch.zeros((n, n), dtype=dtype, device="cuda")
    t_std = ev_time(lambda: standardize_lib(x.clone()))
    xs = standardize_lib(x.clone())
    t_mm = ev_time(lambda: K.add_(xs @ xs.T))
    print("library route, %s: standardize %.1f ms (%.2e genotypes/s), matmul+add %.1f ms (%.1f TFLOP/s, 2N^2M) per %d-SNP block -> %.1f TFLOP/s for standardize+K"
          % (name, t_std, n * block / t_std * 1e3, t_mm, 2.0 * n * n * block / t_mm / 1e9, block, 2.0 * n * n * block / (t_mm + t_std) / 1e9), flush=True)
    del x, xs, K
torch.backends.cuda.matmul.allow_tf32 = False
Kk = torch.zeros((n, n), device="cuda")
t_ours = ev_time(lambda: dev.snp_kernel(store, K=Kk, accumulate=False))
print("this repo, packed bytes -> K (decode + Unit + 2-term fp16 SYRK): %.1f ms per %d-SNP block = %.1f TFLOP/s" % (t_ours, block, 2.0 * n * n * block / t_ours / 1e9))
# decode+standardize on the library route for cfg2's shape (per 100 000-SNP slice): H2D of the float matrix is part of that route
n2, m2 = 10000, 100000
store2 = bench.gen_store_device(dev, torch, n2, m2, seed=2)
raw, _ = dev.read(store2, dtype=np.float32, order="C")
host = raw.cpu().pin_memory()
def lib_route():
    x = host.cuda(non_blocking=True)
    standardize_lib(x)
t = ev_time(lib_route)
print("library route cfg2 slice (H2D float32 matrix + standardize): %.1f ms = %.2e genotypes/s" % (t, n2 * m2 / t * 1e3))
