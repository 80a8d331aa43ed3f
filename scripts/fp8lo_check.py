"""Accuracy of the 2-term SYRK with the low term on the fp8 pipe (PSTB_SYRK_FP8LO=1) against the float64 oracle, data without missing values."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import bed_oracle as o
from pysnptools_b200 import device as dev
for n, m, std, args in ((700, 4096, ("unit",), {}), (2100, 700, ("unit",), {}), (515, 300, ("unit",), {}), (1500, 6000, ("beta", 1, 25), dict(is_beta=True, a=1, b=25))):
    packed = o.synth_packed(n, 0, m, 0.0, seed=n)
    store = dev.PackedStore.from_host(packed, n)
    ref, _ = o.read_kernel(packed, n, **args)
    K, _ = dev.snp_kernel(store, standardizer=std, chunk=min(1024, (m + 63) // 64 * 64))
    Kc = K.double().cpu().numpy()
    print(n, m, std[0], "rel fro %.3e" % (np.linalg.norm(Kc - ref) / np.linalg.norm(ref)), "symmetric", np.array_equal(Kc, Kc.T), flush=True)
