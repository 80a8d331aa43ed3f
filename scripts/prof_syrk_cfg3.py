"""Workload for the round-2 ncu capture of k_syrk2 at the cfg3 shape: N = 50 000, two 4 032-SNP chunks (first launch stores K through
the TMA, the second reduce-adds), fp8 low term; optional argument: missing rate (default 0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pysnptools_b200 import device as dev
missing = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
n, m = 50_000, 2 * 4032
store = bench.gen_store_device(dev, torch, n, m, seed=2000, missing_rate=missing)
K = torch.zeros((n, n), device="cuda")
for _ in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); dev.snp_kernel(store, K=K, accumulate=False, chunk=4032, mirror=False, low_term="fp8"); b.record(); torch.cuda.synchronize()
    print("snp_kernel n=%d m=%d missing=%.2f: %.3f ms  %.1f TFLOP/s (2N^2M)" % (n, m, missing, a.elapsed_time(b), 2.0 * n * n * m / a.elapsed_time(b) / 1e9), flush=True)
