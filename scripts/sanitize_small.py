"""Small cases of every kernel family for compute-sanitizer (memcheck); sizes chosen to touch tails and fallbacks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import bed_oracle as o
from pysnptools_b200 import device as dev
rng = np.random.default_rng(0)
for n, m in ((5, 3), (203, 31), (1003, 17), (4099, 5), (30001, 3)):
    packed = o.synth_packed(n, 0, m, 0.1, seed=n)
    store = dev.PackedStore.from_host(packed, n)
    ii = rng.permutation(n)[: max(1, n // 2)]
    for sel in (None, ii, slice(None, None, -1)):
        for dtype in (np.float32, np.float64, np.int8):
            for order in ("F", "C"):
                v, _ = dev.read(store, sel, None, dtype=dtype, order=order)
                if dtype != np.int8:
                    v, st = dev.read(store, sel, None, dtype=dtype, order=order, standardizer=("beta", 1, 25))
    K, st = dev.snp_kernel(store, chunk=64)
    K, st = dev.snp_kernel(store, ii, None, chunk=64)
    t, c, st = dev.snp_kernel_tiles(store, rank=1, world=3, chunk=64)
    x = torch.randn(n, m, device="cuda", dtype=torch.float64)
    dev.standardize(x, ("unit",)); dev.standardize(x.t().contiguous().t(), ("unit",))
    dev.float_kernel(x.float())
    dev.pack(torch.randint(0, 3, (n, m), device="cuda").to(torch.int8))
    if n <= 4099:
        other = dev.PackedStore.from_host(o.synth_packed(77, 0, m, 0.1, seed=n + 1), 77)
        dev.snp_cross_kernel(store, other, ii, None, None, None, chunk=64)                       # train x test, gathered rows
        from pysnptools_b200 import _lib
        Kh = np.empty((n, n), dtype=np.float64)
        sth = np.empty((m, 2))
        _lib.check(_lib.lib.pstb_snp_kernel_host(packed.ctypes.data, n, m, None, n, None, m, 0, _lib.STD_UNIT, 0.0, 0.0, 0,
                                                 sth.ctypes.data, Kh.ctypes.data, _lib.F64, 64, -1))
# the atomic record feed: more records than resident groups on every F-order kernel (warp-per-record 8x1 and 4x2 shapes, CTA-per-record,
# short records, gathered batches of four, the statistics-only pass behind a C-order read, K2f on a device matrix)
for n, m in ((4104, 2500), (10000, 700), (300, 6000), (30001, 400)):
    packed = o.synth_packed(n, 0, m, 0.05, seed=n + 3)
    store = dev.PackedStore.from_host(packed, n)
    ii = rng.permutation(n)[: n // 2]
    for dtype in (np.float32, np.float64, np.int8):
        dev.read(store, None, None, dtype=dtype, order="F")
    dev.read(store, None, None, dtype=np.float32, order="F", standardizer=("unit",))
    dev.read(store, None, None, dtype=np.float32, order="C", standardizer=("unit",))
    dev.read(store, ii, None, dtype=np.float32, order="F", standardizer=("beta", 1, 25))
    x = torch.randn(n, min(m, 600), device="cuda", dtype=torch.float32).t().contiguous().t()
    dev.standardize(x, ("unit",))
tight = torch.from_numpy(o.synth_packed(203, 0, 9, 0.1, seed=1)).cuda()
dev.read(dev.PackedStore(tight, 203, 9), dtype=np.float32, standardizer=("unit",))
nomiss = o.synth_packed(300, 0, 128, 0.0, seed=2)
dev.snp_kernel(dev.PackedStore.from_host(nomiss, 300), chunk=64)       # 2-term path
torch.cuda.synchronize()
print("sanitize_small done")
