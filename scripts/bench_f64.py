"""The float64 kernel path (syrk_f64.cu) against the tensor-core path on one GPU: time and TFLOP/s (2 N^2 M convention), error vs the float64 oracle on a block."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pysnptools_b200 import device as dev
for n, m in ((10_000, 20_000), (50_000, 10_000)):
    store = bench.gen_store_device(dev, torch, n, m, seed=3, missing_rate=0.02)
    for name, fn in (("float64 path (k_dsyrk)", lambda: dev.snp_kernel_f64(store)), ("tensor-core path (k_syrk2)", lambda: dev.snp_kernel(store))):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); K, st = fn(); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        print("N=%d M=%d %-28s %9.1f ms  %8.1f TFLOP/s (2N^2M; the triangle is half of that in executed flops)" % (n, m, name, ms, 2.0 * n * n * m / ms / 1e9), flush=True)
        del K
    del store
