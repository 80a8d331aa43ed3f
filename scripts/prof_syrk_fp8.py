"""Workload for the ncu capture of k_syrk2 with the low term on the fp8 pipe: n = 16 384, m = 32 768, no missing data, 8 192-SNP chunks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pysnptools_b200 import device as dev
n, m = 16384, 32768
rng = np.random.default_rng(0)
packed = rng.integers(0, 256, size=(m, (n + 3) // 4), dtype=np.uint8)
packed &= ~((packed & 0x55) & ~((packed >> 1) & 0x55))   # turn code 01 (missing) into 00
store = dev.PackedStore.from_host(packed, n)
K = torch.zeros((n, n), device="cuda")
for _ in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); dev.snp_kernel(store, K=K, accumulate=False, chunk=8192); b.record(); torch.cuda.synchronize()
    print("snp_kernel n=%d m=%d low term %s: %.3f ms  %.1f TFLOP/s (2N^2M)" % (n, m, dev.get_syrk_low_term(), a.elapsed_time(b), 2.0 * n * n * m / a.elapsed_time(b) / 1e9), flush=True)
