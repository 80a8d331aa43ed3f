"""Workload for the ncu captures of the round-1 late kernels: k_stats_dense + k_emit_c_wide (C-order read + Unit) and k_std_f_staged (K2f)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pysnptools_b200 import device as dev
n, m = 10000, 200000
t = torch.randint(0, 256, (m, 2512), dtype=torch.uint8, device="cuda")
store = dev.PackedStore(t, n, m)
for _ in range(2):
    val, st = dev.read(store, dtype=np.float32, order="C", standardizer=("unit",))
del val
x = torch.randint(0, 3, (m, n), device="cuda").to(torch.float32).t()          # F order [n, m]
for _ in range(2):
    dev.standardize(x, ("unit",))
torch.cuda.synchronize()
print("ok")
