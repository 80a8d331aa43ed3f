"""One GPU, no collective: the band-major tail of parallel.snp_kernel_sharded_overlapped against the plain chunk-major kernel (experiment)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pysnptools_b200 import device as dev, parallel
n, m = 50_000, 62_500
store = bench.gen_store_device(dev, torch, n, m, seed=2000)
tiles = torch.zeros((len(dev.kernel_tile_coords(n)), 256, 256), dtype=torch.float32, device="cuda")
K = torch.zeros((n, n), dtype=torch.float32, device="cuda")
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
print("plain snp_kernel_tiles + kernel_from_tiles: %.1f ms" % timeit(lambda: (dev.snp_kernel_tiles(store, tiles=tiles, accumulate=False, low_term="fp8"), dev.kernel_from_tiles(tiles, n, K=K))), flush=True)
for tail, bands in ((4, 8), (4, 1), (1, 8), (1, 1), (2, 4), (4, 2)):
    t = timeit(lambda: parallel.snp_kernel_sharded_overlapped(store, n, 500_000, None, ("unit",), tiles=tiles, K=K, bands=bands, tail_chunks=tail))
    print("band-major tail: %d chunks x %d bands: %.1f ms" % (tail, bands, t), flush=True)
