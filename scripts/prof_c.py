import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pysnptools_b200 import device as dev
n, m = 10000, 200000
t = torch.randint(0, 256, (m, 2512), dtype=torch.uint8, device="cuda")
store = dev.PackedStore(t, n, m)
for _ in range(3):
    val, st = dev.read(store, dtype=np.float32, order="C", standardizer=("unit",))
torch.cuda.synchronize()
print("ok", val.shape)
