"""cfg1 (tests/datasets all_chr.maf0.001.N300, 300 x 1015): wall time of the reference-facing calls vs the CPU oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pysnptools_b200 import Bed, Unit
from oracle import bed_oracle
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "data", "n300.bed")
bed = Bed(path, count_A1=False)
bed.iid, bed.sid, bed.pos
def t(fn, reps=20):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e3
print("Bed.read(f64) + Unit().standardize (2 calls): %.3f ms" % t(lambda: bed.read().standardize(Unit())))
print("Bed.read(f64, standardizer=Unit()) fused    : %.3f ms" % t(lambda: bed.read(standardizer=Unit())))
print("Bed.read_kernel(Unit())                     : %.3f ms" % t(lambda: bed.read_kernel(Unit())))
packed = bed_oracle.read_packed(path, 300, 1015)
print("CPU oracle (NumPy) decode + Unit            : %.3f ms" % t(lambda: bed_oracle.standardize(bed_oracle.decode(packed, 300)), reps=5))
print("CPU oracle (NumPy) read_kernel              : %.3f ms" % t(lambda: bed_oracle.read_kernel(packed, 300), reps=5))
