"""Micro-benchmark of the read kernels (K1/K2) on one B200: cfg2 / cfg4 shapes, dtypes, orders."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pysnptools_b200 import _lib, device as dev
lib = _lib.lib
PEAK = 6550.7

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))

def rand_store(n, m, missing=False):
    ld = int(lib.pstb_packed_ld(n))
    if os.environ.get("PSTB_BENCH_LD_ALIGN"):
        al = int(os.environ["PSTB_BENCH_LD_ALIGN"])
        ld = ((n + 3) // 4 + al - 1) // al * al
    t = torch.randint(0, 256, (m, ld), dtype=torch.uint8, device="cuda")
    if not missing:
        lo = t & 0x55; hi = (t >> 1) & 0x55
        t = t & ~(lo & ~hi)            # code 01 -> 00
    return dev.PackedStore(t, n, m)

def run(tag, n, m, dtype, order, std, isel=None, ssel=None, missing=False, store=None):
    store = store or rand_store(n, m, missing)
    I = dev.Selection(isel, n, "cuda"); S = dev.Selection(ssel, m, "cuda")
    es = np.dtype(dtype).itemsize
    code, base, view = dev._alloc_out(I.n, S.n, dtype, order, "cuda")
    stats = torch.empty((S.n, 2), dtype=torch.float64, device="cuda")
    mode, a, b = dev._mode_args(std)
    st = torch.cuda.current_stream().cuda_stream
    fn = lambda: _lib.check(lib.pstb_decode_standardize(store.tensor.data_ptr(), store.ld, n, m, I.axis(), S.axis(), 0, mode, a, b, 0,
                                                        stats.data_ptr(), base.data_ptr(), code, dev._order_code(order), st))
    ms = timeit(fn)
    algo = S.n * ((n + 3) // 4) + es * I.n * S.n + (16 * S.n if std else 0)
    print("%-34s n=%d m=%d out=%dx%d %s %s: %.3f ms  %.3e genotypes/s  %.0f GB/s algorithmic = %.1f%% of %.0f"
          % (tag, n, m, I.n, S.n, np.dtype(dtype).name, order, ms, I.n * S.n / ms * 1e3, algo / ms / 1e6, 100 * algo / ms / 1e6 / PEAK, PEAK), flush=True)
    del base, view

import signal; signal.signal(signal.SIGPIPE, signal.SIG_DFL)
which = sys.argv[1:] or ["cfg2", "cfg4", "variants"]
if "peaks" in which:
    a = torch.empty(10_000_000_000 // 4, dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
    ms = timeit(lambda: a.fill_(1.0)); print("torch fill_ 10 GB (write only): %.3f ms  %.0f GB/s" % (ms, 10e9 / ms / 1e6))
    ms = timeit(lambda: b.copy_(a)); print("torch copy_ 10 GB->10 GB: %.3f ms  %.0f GB/s (read+write)" % (ms, 20e9 / ms / 1e6))
    ms = timeit(lambda: a.sum()); print("torch sum 10 GB (read only): %.3f ms  %.0f GB/s" % (ms, 10e9 / ms / 1e6), flush=True)
    del a, b
if "cfg2" in which:
    run("cfg2 decode+Unit", 10000, 1000000, np.float32, "F", ("unit",))
if "cfg4real" in which:
    # the same selection on a store with the genotype distribution of SURVEY 8d (p ~ U(.05,.5), Binomial(2,p), 5 % missing) instead of random bytes
    import bench
    rng = np.random.default_rng(1)
    N, M = 100000, 200000
    ii, si = rng.permutation(N)[: N // 2], rng.permutation(M)[: M // 2]
    real = bench.gen_store_device(dev, torch, N, M, seed=4000, missing_rate=0.05)
    run("cfg4 real genotypes, Beta", N, M, np.float32, "F", ("beta", 1, 25), ii, si, store=real)
    run("cfg4 real genotypes, Unit", N, M, np.float32, "F", ("unit",), ii, si, store=real)
    run("cfg4 real genotypes, decode only", N, M, np.float32, "F", None, ii, si, store=real)
    run("cfg4 real, sorted iid subset", N, M, np.float32, "F", ("beta", 1, 25), np.sort(ii), si, store=real)
    run("cfg4 random bytes, Beta", N, M, np.float32, "F", ("beta", 1, 25), ii, si, missing=True)
    run("cfg4 random bytes, Unit", N, M, np.float32, "F", ("unit",), ii, si, missing=True)
if "cfg4" in which:
    rng = np.random.default_rng(1)
    N, M = 100000, 200000
    run("cfg4 Beta gather 1/2 x 1/2", N, M, np.float32, "F", ("beta", 1, 25), rng.permutation(N)[: N // 2], rng.permutation(M)[: M // 2], missing=True)
    run("cfg4-like dense full", N, M // 2, np.float32, "F", ("beta", 1, 25), missing=True)
if "variants" in which:
    run("decode only f32 F", 10000, 500000, np.float32, "F", None)
    run("decode+Unit f64 F", 10000, 400000, np.float64, "F", ("unit",))
    run("decode i8 F", 10000, 1000000, np.int8, "F", None)
    run("decode+Unit f32 C", 10000, 500000, np.float32, "C", ("unit",))
    run("decode+Unit f64 C", 10000, 250000, np.float64, "C", ("unit",))
    run("decode+Unit f32 F N=50k (cta)", 50000, 200000, np.float32, "F", ("unit",))
    run("decode+Unit f32 F N=500k (cta)", 500000, 20000, np.float32, "F", ("unit",))
    run("decode+Unit f32 F N=300", 300, 4000000, np.float32, "F", ("unit",))
if "stats" in which:
    # statistics only (no output): the first pass of a C-order read and of every K3 chunk
    for n, m in ((10000, 500000), (50000, 100000)):
        store = rand_store(n, m)
        stats = torch.empty((m, 2), dtype=torch.float64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        fn = lambda: _lib.check(lib.pstb_decode_standardize(store.tensor.data_ptr(), store.ld, n, m, _lib.Axis(None, 0, 1, n), _lib.Axis(None, 0, 1, m), 0,
                                                            _lib.STD_UNIT, float("nan"), float("nan"), 0, stats.data_ptr(), None, _lib.F32, _lib.ORDER_F, st))
        ms = timeit(fn)
        print("statistics only n=%d m=%d: %.3f ms  %.0f GB/s of packed reads" % (n, m, ms, m * ((n + 3) // 4) / ms / 1e6), flush=True)
        del store
if "k2f" in which:
    # K2f: standardize an existing float matrix in place (read + write = 2 * esize bytes per value)
    for dt, es in ((torch.float32, 4), (torch.float64, 8)):
        for n, m in ((10000, 400000 if es == 4 else 200000), (50000, 80000 if es == 4 else 40000)):
            base = torch.randint(0, 3, (m, n), device="cuda").to(dt)
            base[::97, ::89] = float("nan")
            for order, v in (("F", base.t()), ("C", base.t().contiguous())):
                src = v.clone() if order == "C" else None
                def fn():
                    dev.standardize(v, ("unit",))
                ms = timeit(fn, reps=3, warm=1)
                print("K2f Unit %s %s n=%d m=%d: %.3f ms  %.0f GB/s (read + write)" % (str(dt), order, n, m, ms, 2 * es * n * m / ms / 1e6), flush=True)
            del base, v
    # sub_matrix gather: half the rows x half the columns, random
    for dt, es in ((torch.float32, 4), (torch.float64, 8)):
        n, m = 20000, 100000 if es == 4 else 50000
        src = torch.randn((n, m), device="cuda", dtype=dt)
        rng = np.random.default_rng(0)
        rows = dev.Selection(rng.permutation(n)[: n // 2], n, "cuda"); cols = dev.Selection(rng.permutation(m)[: m // 2], m, "cuda")
        out = torch.empty((rows.n, cols.n), device="cuda", dtype=dt)
        code = _lib.F32 if es == 4 else _lib.F64
        st = torch.cuda.current_stream().cuda_stream
        fn = lambda: _lib.check(lib.pstb_subset(src.data_ptr(), code, _lib.ORDER_C, n, m, 1, rows.axis(), cols.axis(), out.data_ptr(), code, _lib.ORDER_C, st))
        ms = timeit(fn)
        print("subset %s C->C %dx%d of %dx%d: %.3f ms  %.0f GB/s (out bytes x 2)" % (str(dt), rows.n, cols.n, n, m, ms, 2 * es * rows.n * cols.n / ms / 1e6), flush=True)
        del src, out
if "pack" in which:
    for dt, es in ((torch.int8, 1), (torch.float32, 4)):
        n, m = 10000, 200000
        v = torch.randint(0, 3, (m, n), device="cuda").to(dt).t()       # F-order [n, m]
        ms = timeit(lambda: dev.pack(v))
        print("pack %s F n=%d m=%d: %.3f ms  %.0f GB/s (read %d B + write 0.25 B per genotype)" % (str(dt), n, m, ms, n * m * (es + 0.25) / ms / 1e6, es), flush=True)
        vc = v.contiguous()
        ms = timeit(lambda: dev.pack(vc))
        print("pack %s C n=%d m=%d: %.3f ms  %.0f GB/s" % (str(dt), n, m, ms, n * m * (es + 0.25) / ms / 1e6), flush=True)
if "nfine" in which:
    # does the F-order write efficiency depend on the column stride (DRAM channel camping)?
    for n in (8192, 9984, 10000, 10016, 10240, 10496, 11264, 12288, 16384):
        m = int(8e9 // (4 * n)) // 8 * 8
        run("decode+Unit f32 F N=%d" % n, n, m, np.float32, "F", ("unit",))
if "nmid" in which:
    for n in (2400, 3000, 4000, 6000):
        m = int(4e9 // (4 * n)) // 8 * 8
        run("decode+Unit f32 F N=%d" % n, n, m, np.float32, "F", ("unit",))
if "nsweep" in which:
    for n in (300, 1000, 2000, 4000, 8000, 16000, 24000, 32000):
        m = int(4e9 // (4 * n)) // 8 * 8
        run("decode+Unit f32 F N=%d" % n, n, m, np.float32, "F", ("unit",))
if "feed" in which:
    # k_read_f<warp> knobs (read per launch) under the dynamic record feed: warps per CTA x CTAs per SM, on cfg2 and neighbouring shapes
    st2 = rand_store(10000, 1000000)
    for warps, ctas in ((8, 1), (8, 2), (6, 1), (6, 2), (4, 1), (4, 2), (4, 3), (4, 4), (2, 2), (2, 4)):
        os.environ["PSTB_READ_WARPS"] = str(warps); os.environ["PSTB_READ_CTAS"] = str(ctas)
        run("cfg2 warps=%d ctas=%d" % (warps, ctas), 10000, 1000000, np.float32, "F", ("unit",), store=st2)
    for warps, ctas in ((8, 1), (8, 2), (4, 2)):
        os.environ["PSTB_READ_WARPS"] = str(warps); os.environ["PSTB_READ_CTAS"] = str(ctas)
        run("cfg2 f64 warps=%d ctas=%d" % (warps, ctas), 10000, 500000, np.float64, "F", ("unit",), store=dev.PackedStore(st2.tensor[:500000], 10000, 500000))
        run("cfg2 decode only warps=%d ctas=%d" % (warps, ctas), 10000, 1000000, np.float32, "F", None, store=st2)
    del st2
    for n in (4104, 6000, 20000):
        stn = rand_store(n, int(1e10 // n // 8 * 8))
        for warps, ctas in ((8, 1), (8, 2), (4, 2)):
            os.environ["PSTB_READ_WARPS"] = str(warps); os.environ["PSTB_READ_CTAS"] = str(ctas)
            run("warps=%d ctas=%d" % (warps, ctas), n, stn.sid_count, np.float32, "F", ("unit",), store=stn)
        del stn
    del os.environ["PSTB_READ_WARPS"], os.environ["PSTB_READ_CTAS"]
if "feed_i8" in which:
    st2 = rand_store(10000, 1000000)
    for warps, ctas in ((8, 1), (8, 2), (8, 3), (8, 4), (4, 2), (4, 4), (4, 6)):
        os.environ["PSTB_READ_WARPS"] = str(warps); os.environ["PSTB_READ_CTAS"] = str(ctas)
        run("cfg2 int8 decode warps=%d ctas=%d" % (warps, ctas), 10000, 1000000, np.int8, "F", None, store=st2)
    del os.environ["PSTB_READ_WARPS"], os.environ["PSTB_READ_CTAS"]
    del st2
