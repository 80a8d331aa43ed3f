"""Micro-benchmark of the K3 stages on one B200 (not the contract bench; see bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pysnptools_b200 import device as dev
from pysnptools_b200._lib import lib, check

def timeit(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
k = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
n_pad = (n + 255) // 256 * 256
x = torch.randn((n_pad, k), device="cuda") * 2
x[n:] = 0
hi = x.half(); lo = (x - hi.float()).half()
K = torch.zeros((n, n), device="cuda")
st = torch.cuda.current_stream().cuda_stream
fn = lambda: check(lib.pstb_syrk_planes(hi.data_ptr(), lo.data_ptr(), n, n_pad, k, K.data_ptr(), n, 0, 1.0, st))
best, med = timeit(fn)
ntiles = sum(1 for I in range((n + 127) // 128) for J in range((n + 255) // 256) if J * 256 <= I * 128 + 127)
exec_flops = ntiles * 3 * 2 * 128 * 256 * k
print("syrk_planes n=%d k=%d: best %.3f ms median %.3f ms | executed %.1f TFLOP/s | algorithmic(2N^2K) %.1f TFLOP/s"
      % (n, k, best, med, exec_flops / best / 1e9, 2.0 * n * n * k / best / 1e9))
# accuracy of the stage vs fp64 on a corner
sub = slice(0, 512)
ref = (hi[sub].double() + lo[sub].double()) @ (hi[sub].double() + lo[sub].double()).T
got = torch.tril(K[sub, sub].double()); ref = torch.tril(ref)
print("rel fro (512x512 corner):", float(torch.linalg.norm(got - ref) / torch.linalg.norm(ref)),
      " diag rel max:", float(((got.diagonal() - ref.diagonal()).abs() / ref.diagonal()).max()))
# end-to-end snp_kernel on a synthetic store
m = k
rng = np.random.default_rng(0)
packed = rng.integers(0, 256, size=(m, (n + 3) // 4), dtype=np.uint8)
packed &= ~((packed & 0x55) & ~((packed >> 1) & 0x55))   # turn code 01 (missing) into 00
store = dev.PackedStore.from_host(packed, n)
Kk = torch.zeros((n, n), device="cuda")
fn2 = lambda: dev.snp_kernel(store, K=Kk, accumulate=False, chunk=min(m, 8192))
best2, med2 = timeit(fn2)
print("snp_kernel n=%d m=%d: best %.3f ms | algorithmic %.1f TFLOP/s" % (n, m, best2, 2.0 * n * n * m / best2 / 1e9))
