#!/usr/bin/env python
"""bench.py -- the genotype hot path on B200: decode 2-bit .bed -> Unit standardize (float32), and SnpKernel.

Contract line (one JSON object on stdout, rank 0):
  metric  "genotypes/s decoded+standardized" on BASELINE.json configs[1]
          (synthetic .bed, 10 000 iids x 1 000 000 SNPs, Unit, float32, F order, 1 B200 per rank).
  value   whole-job throughput with the packed bytes resident in HBM (one fused kernel launch per step).
  e2e     the same workload through the host-buffer C ABI (pstb_read_host: pinned host packed bytes in,
          pinned host float32 matrix out, H2D and D2H inside the timed region).
  roofline / cpu_baseline / clocks / gpu_launches as the build brief defines them.
  kernel  (extra object) SnpKernel TFLOP/s on configs[2] (50 000 x 500 000, Unit, K fp32), 2*N^2*M convention.

`--impl reference` times the reference's CPU path for the same metric: the plain-C/OpenMP port in oracle/
(the reference's own native code is the external Rust wheel `bed-reader`, not in this tree and not
buildable here -- DESIGN.md) on a bounded sample of the same workload, all host threads.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG2 = dict(n_iid=10_000, n_sid=1_000_000)
CFG3 = dict(n_iid=50_000, n_sid=500_000)


# ------------------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port)
# ------------------------------------------------------------------------------------------------------------
def _oracle_lib():
    so = os.path.join(ROOT, "oracle", "_build", "libpst_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return ctypes.CDLL(so)


def synth_packed_host(n_iid, n_sid, seed=0):
    """Synthetic packed records on the host (same distribution as the device generator)."""
    rng = np.random.default_rng(seed)
    rec = (n_iid + 3) // 4
    out = np.empty((n_sid, rec), dtype=np.uint8)
    code_of = np.array([0, 2, 3], dtype=np.uint8)
    step = max(1, (64 << 20) // max(1, n_iid))
    for s0 in range(0, n_sid, step):
        s1 = min(n_sid, s0 + step)
        p = rng.uniform(0.05, 0.5, size=(s1 - s0, 1)).astype(np.float32)
        g = (rng.random((s1 - s0, n_iid), dtype=np.float32) < p).astype(np.uint8) + (rng.random((s1 - s0, n_iid), dtype=np.float32) < p)
        c = np.zeros((s1 - s0, rec * 4), dtype=np.uint8)
        c[:, :n_iid] = code_of[g]
        c = c.reshape(s1 - s0, rec, 4)
        out[s0:s1] = c[:, :, 0] | (c[:, :, 1] << 2) | (c[:, :, 2] << 4) | (c[:, :, 3] << 6)
    return out


def cpu_decode_standardize(lib, packed, n_iid, threads, reps=1):
    """Reference CPU path: decode to a float32 F-order matrix, then standardize it in place (Unit)."""
    n_sid = packed.shape[0]
    out = np.empty((n_iid, n_sid), dtype=np.float32, order="F")
    stats = np.empty((n_sid, 2), dtype=np.float64)
    p = ctypes.c_void_p
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        lib.pst_oracle_decode_f32(p(packed.ctypes.data), ctypes.c_int64(packed.shape[1]), ctypes.c_int64(n_iid), ctypes.c_int64(n_sid),
                                  None, ctypes.c_int64(n_iid), None, ctypes.c_int64(n_sid), 0, 0, p(out.ctypes.data), threads)
        lib.pst_oracle_standardize_f32(p(out.ctypes.data), ctypes.c_int64(n_iid), ctypes.c_int64(n_sid), 0, 0, ctypes.c_double(np.nan),
                                       ctypes.c_double(np.nan), 0, p(stats.ctypes.data), threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, out, stats


def cpu_snp_kernel(lib, n_iid, sample_sid, threads, dtype, seed=0):
    """Reference CPU path of SnpReader._read_kernel (snpreader.py:651-655) for ONE block of `sample_sid` SNPs at full N:
    decode -> standardize (Unit) -> K += val.dot(val.T) with NumPy's BLAS.  Returns (seconds, TFLOP/s in the 2*N^2*M convention)."""
    packed = synth_packed_host(n_iid, sample_sid, seed=seed)
    K = np.zeros((n_iid, n_iid), dtype=dtype)
    t0 = time.perf_counter()
    _, val32, _ = cpu_decode_standardize(lib, packed, n_iid, threads)
    val = val32 if dtype == np.float32 else val32.astype(np.float64)        # the reference reads in the kernel's dtype (snpreader.py:652)
    K += val.dot(val.T)
    dt = time.perf_counter() - t0
    del K, val, val32
    return dt, 2.0 * n_iid * n_iid * sample_sid / dt / 1e12


def cpu_kernel_baseline(lib, n_iid, sample_sid, threads):
    # torchrun exports OMP_NUM_THREADS=1; the CPU baseline gets every host core, as the reference would use them
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=threads)
    except Exception:
        limiter = None
    try:
        t32, f32 = cpu_snp_kernel(lib, n_iid, sample_sid, threads, np.float32)
        t64, f64 = cpu_snp_kernel(lib, n_iid, max(64, sample_sid // 2), threads, np.float64)
    finally:
        if limiter is not None:
            limiter.restore_original_limits()
    return {"value": f32, "unit": "TFLOP/s", "cores": threads, "kind": "port",
            "sample": "one block of {0} SNPs x {1} iids of the cfg3 workload: oracle/c decode + Unit standardize, then NumPy val.dot(val.T) (BLAS, all cores) "
                      "into a float32 K; seconds = {2:.2f}".format(sample_sid, n_iid, t32),
            "float64_value": f64, "float64_note": "the reference's default dtype (K float64), block of {0} SNPs, {1:.2f} s".format(max(64, sample_sid // 2), t64)}


class limit_threads(object):
    """torchrun exports OMP_NUM_THREADS=1; the CPU checker / baseline legs get their share of the host cores back."""

    def __init__(self, threads):
        self.threads, self.limiter = threads, None

    def __enter__(self):
        try:
            from threadpoolctl import threadpool_limits
            self.limiter = threadpool_limits(limits=self.threads)
        except Exception:
            self.limiter = None
        return self

    def __exit__(self, *exc):
        if self.limiter is not None:
            self.limiter.restore_original_limits()
        return False


def oracle_rows(lib, store, stats, n, block, spec, threads):
    """The CPU oracle's standardized values of the 256 individuals of row block `block` for every SNP of this rank's store:
    float64 [cnt, m_local] (C order).  256 aligned individuals are 64 contiguous bytes of every packed record."""
    lo_b = block * 64
    cnt = min(n, block * 256 + 256) - block * 256
    sub = np.ascontiguousarray(store.tensor[:, lo_b:lo_b + (cnt + 3) // 4].contiguous().cpu().numpy())
    m_local = sub.shape[0]
    x = np.empty((cnt, m_local), dtype=np.float64)
    p = ctypes.c_void_p
    lib.pst_oracle_decode_f64(p(sub.ctypes.data), ctypes.c_int64(sub.shape[1]), ctypes.c_int64(cnt), ctypes.c_int64(m_local), None, ctypes.c_int64(cnt),
                              None, ctypes.c_int64(m_local), 0, 1, p(x.ctypes.data), threads)
    is_beta = spec[0] == "beta"
    a, b = (float(spec[1]), float(spec[2])) if is_beta else (float("nan"), float("nan"))
    lib.pst_oracle_standardize_f64(p(x.ctypes.data), ctypes.c_int64(cnt), ctypes.c_int64(m_local), 1, int(is_beta), ctypes.c_double(a), ctypes.c_double(b),
                                   1, p(stats.ctypes.data), threads)
    return x


def sampled_tile_parity(torch, dist, world, rank, lib_o, store, stats_t, n, spec, fetch_tile, blocks, owner_of=None):
    """K against the CPU oracle on sampled 256 x 256 tiles: every pair (I, J), I >= J, of the row blocks in `blocks` (diagonal tiles
    included), over ALL SNPs of the kernel.  Every rank builds the float64 reference of its own SNP shard (oracle/c decode +
    standardize with the statistics of the GPU run -- themselves checked against the oracle on the shard's first SNPs at full N --
    then a NumPy float64 GEMM); the partial references are summed over the ranks.  fetch_tile(I, J) returns the GPU result tile as a
    float64 array, or None when another rank holds it (K-tile sharding: owner_of(I, J) says which).
    Returns per-tile relative Frobenius errors, a stratified estimate for the whole matrix, and the relative bias of the diagonal."""
    threads = max(1, (os.cpu_count() or 1) // max(1, world))
    stats = np.ascontiguousarray(stats_t.cpu().numpy())
    m_local = stats.shape[0]
    rec = (n + 3) // 4
    head = min(64, m_local)
    if head:
        packed_head = np.ascontiguousarray(store.tensor[:head, :rec].contiguous().cpu().numpy())
        full = np.empty((n, head), dtype=np.float64, order="F")
        st_ref = np.empty((head, 2), dtype=np.float64)
        p = ctypes.c_void_p
        lib_o.pst_oracle_decode_f64(p(packed_head.ctypes.data), ctypes.c_int64(rec), ctypes.c_int64(n), ctypes.c_int64(head), None, ctypes.c_int64(n),
                                    None, ctypes.c_int64(head), 0, 0, p(full.ctypes.data), threads)
        is_beta = spec[0] == "beta"
        a, b = (float(spec[1]), float(spec[2])) if is_beta else (float("nan"), float("nan"))
        lib_o.pst_oracle_standardize_f64(p(full.ctypes.data), ctypes.c_int64(n), ctypes.c_int64(head), 0, int(is_beta), ctypes.c_double(a), ctypes.c_double(b),
                                         0, p(st_ref.ctypes.data), threads)
        stats_ok = bool(np.allclose(stats[:head], st_ref, rtol=1e-12, equal_nan=True))
        del full
    else:
        stats_ok = True
    pairs = [(I, J) for a_, I in enumerate(blocks) for J in blocks[: a_ + 1]]
    pairs = [(max(I, J), min(I, J)) for I, J in pairs]
    with limit_threads(threads):
        rows = {b: oracle_rows(lib_o, store, stats, n, b, spec, threads) for b in blocks}
        ref = np.zeros((len(pairs), 256, 256), dtype=np.float64)
        for t, (I, J) in enumerate(pairs):
            r = rows[I].dot(rows[J].T)
            ref[t, : r.shape[0], : r.shape[1]] = r
    del rows
    if world > 1:
        ref_t = torch.from_numpy(ref).cuda()
        dist.all_reduce(ref_t)
        ok_t = torch.tensor([1.0 if stats_ok else 0.0], device="cuda")
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
        stats_ok = bool(ok_t.item() > 0.5)
        ref = ref_t.cpu().numpy()
        del ref_t
    # per-tile errors where the tile lives; with K-tile sharding the partial sums are combined over the ranks
    acc = np.zeros((len(pairs), 4), dtype=np.float64)                   # diff^2, ref^2, sum of relative diagonal error, diagonal count
    for t, (I, J) in enumerate(pairs):
        got = fetch_tile(I, J)
        if got is None:
            continue
        ri, rj = min(n, I * 256 + 256) - I * 256, min(n, J * 256 + 256) - J * 256
        d = got[:ri, :rj] - ref[t, :ri, :rj]
        acc[t, 0], acc[t, 1] = float((d * d).sum()), float((ref[t, :ri, :rj] ** 2).sum())
        if I == J:
            dg = np.diagonal(ref[t, :ri, :ri])
            acc[t, 2], acc[t, 3] = float((np.diagonal(d) / np.where(dg != 0, dg, 1.0)).sum()), ri
    if world > 1 and owner_of is not None:
        acc_t = torch.from_numpy(acc).cuda()
        dist.all_reduce(acc_t)
        acc = acc_t.cpu().numpy()
    per_tile = np.sqrt(acc[:, 0] / np.maximum(acc[:, 1], 1e-300))
    diag = np.array([I == J for I, J in pairs])
    T = (n + 255) // 256
    # whole-matrix estimate: T diagonal tiles + T (T - 1) off-diagonal ones (both triangles), each stratum scaled from its sample
    def stratum(mask, population):
        k = int(mask.sum())
        return (acc[mask, 0].sum() * population / k, acc[mask, 1].sum() * population / k) if k else (0.0, 0.0)
    d2a, r2a = stratum(diag, T)
    d2b, r2b = stratum(~diag, T * (T - 1))
    whole = float(np.sqrt((d2a + d2b) / max(r2a + r2b, 1e-300)))
    return {"sampled_tiles": len(pairs), "row_blocks": [int(b) for b in blocks], "diagonal_tiles": int(diag.sum()),
            "worst_rel_frobenius_vs_oracle": float(per_tile.max()),
            "worst_rel_frobenius_diagonal_tile": float(per_tile[diag].max()) if diag.any() else None,
            "worst_rel_frobenius_off_diagonal_tile": float(per_tile[~diag].max()) if (~diag).any() else None,
            "rel_frobenius_whole_K_estimate": whole,
            "diag_rel_bias": float(acc[diag, 2].sum() / max(1.0, acc[diag, 3].sum())) if diag.any() else None,
            "stats_match_oracle_rtol_1e-12": stats_ok, "gate": 1e-5,
            "oracle": "oracle/c decode + standardize (float64, the GPU run's statistics, checked on each shard's first 64 SNPs at full N) + NumPy float64 GEMM over ALL SNPs"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lib = _oracle_lib()
    threads = os.cpu_count() or 1
    n_iid, sample_sid = CFG2["n_iid"], args.ref_sample_sid
    packed = synth_packed_host(n_iid, sample_sid, seed=0)
    for _ in range(args.warmup):
        cpu_decode_standardize(lib, packed, n_iid, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_decode_standardize(lib, packed, n_iid, threads)
    dt = (time.perf_counter() - t0) / args.steps
    value = n_iid * sample_sid / dt
    sample = "{0} iids x {1} SNPs (first {1} of the 1 000 000 SNP workload) per step".format(n_iid, sample_sid)
    line = {
        "impl": "reference", "metric": "genotypes/s decoded+standardized", "value": value, "unit": "genotypes/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: synthetic .bed 10000 iids x 1000000 SNPs, decode + Unit standardize float32, F order",
                   "reference_arm": "oracle/c/pst_oracle.c (plain C + OpenMP port of the reference CPU path; bed-reader Rust wheel absent)"},
        "cpu_baseline": {"value": value, "unit": "genotypes/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "genotypes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.kernel:
        kb = cpu_kernel_baseline(lib, args.kernel_n, args.ref_kernel_sid, threads)
        line["kernel"] = {"metric": "SnpKernel TFLOP/s (2*N^2*M)", "value": kb["value"], "unit": "TFLOP/s", "n_gpus": args.gpus,
                          "config": {"workload": "cfg3: synthetic .bed {0} iids x {1} SNPs, SnpKernel(Unit), K fp32".format(args.kernel_n, args.kernel_m)},
                          "cpu_baseline": kb}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons (B200_PROFILING.md recipe), sampled every 20 ms for the whole run; windows are
    cut out afterwards by wall-clock time (nvidia-smi needs a few hundred ms to start, the cfg2 timed region lasts ~80 ms)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.thread = [], None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def wait_started(self, timeout=3.0):
        t_end = time.perf_counter() + timeout
        while self.proc is not None and not self.rows and time.perf_counter() < t_end:
            time.sleep(0.02)

    def window(self, t0, t1):
        """Median SM clock and the throttle reasons seen between t0 and t1 (the nearest later samples if the window is shorter
        than the sampling period)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        t_end = time.perf_counter() + 0.5
        while time.perf_counter() < t_end and not any(t >= t1 for t, _ in self.rows):
            time.sleep(0.02)
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.03]
        if not rows:
            rows = [r for (t, r) in self.rows if t > t1][:2] or [r for (_, r) in self.rows[-2:]]
        sm, smax, power, reasons = [], None, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
                power.append(float(f[2]))
                for k, nm in enumerate(names):
                    if f[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "power_w_max": max(power) if power else None,
                "reasons": sorted(reasons), "samples": len(sm)}

    def close(self):
        if self.proc is not None:
            self.proc.terminate()


def gen_store_device(dev, torch, n_iid, n_sid, seed, missing_rate=0.0):
    """Synthetic packed store generated on the GPU and packed with the library's own pack kernel (K0)."""
    ld = int(dev.lib.pstb_packed_ld(n_iid))
    store_t = torch.zeros((n_sid, ld), dtype=torch.uint8, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(seed)
    step = max(1, min(n_sid, (1 << 28) // max(1, n_iid)))
    for s0 in range(0, n_sid, step):
        s1 = min(n_sid, s0 + step)
        p = torch.empty((s1 - s0, 1), device="cuda").uniform_(0.05, 0.5, generator=g)
        val = (torch.rand((s1 - s0, n_iid), device="cuda", generator=g) < p).to(torch.int8)
        val += (torch.rand((s1 - s0, n_iid), device="cuda", generator=g) < p).to(torch.int8)
        if missing_rate > 0:
            val[torch.rand((s1 - s0, n_iid), device="cuda", generator=g) < missing_rate] = -127
        part = dev.pack(val.t())                                    # [n_iid, chunk] F-order view of the [chunk, n_iid] buffer
        store_t[s0:s1].copy_(part.tensor)
        del val, part
    return dev.PackedStore(store_t, n_iid, n_sid)


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from pysnptools_b200 import _lib, device as dev, parallel
    lib = _lib.lib
    globals()["parallel"] = parallel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa_node = None
    if world > 1 and args.numa_bind:
        _lib.require_gpu()
        numa_node = parallel.bind_to_gpu_numa_node(local)         # before NCCL and the pinned buffers: both then live next to the GPU
    globals()["NUMA_NODE"] = numa_node
    if world > 1:
        # NCCL announces its version on stdout when the communicator is created; the contract is ONE JSON line on stdout,
        # so stdout points at stderr until the first collective has completed
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    _lib.require_gpu()
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.cfg5:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        res = run_cfg5(args, torch, dist, dev, _lib, rank, world, barrier, max_over_ranks, peaks)
        if rank == 0:
            print(json.dumps(res))
        if world > 1:
            dist.destroy_process_group()
        return
    if args.only_kernel:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        kernel = run_kernel_workload(args, torch, dist, dev, _lib, rank, world, barrier, max_over_ranks, peaks)
        if rank == 0:
            print(json.dumps(kernel))
        if world > 1:
            dist.destroy_process_group()
        return

    n_iid, n_sid = CFG2["n_iid"], CFG2["n_sid"]
    rec = (n_iid + 3) // 4
    # every rank owns one full cfg2-sized SNP shard (weak scaling; no data-path collective -- DESIGN.md)
    store = gen_store_device(dev, torch, n_iid, n_sid, seed=1000 + rank)
    out = torch.empty((n_sid, n_iid), dtype=torch.float32, device="cuda")          # F order: [n_iid, n_sid] transposed
    stats = torch.empty((n_sid, 2), dtype=torch.float64, device="cuda")
    full = _lib.Axis(None, 0, 1, n_iid), _lib.Axis(None, 0, 1, n_sid)

    def step():
        _lib.check(lib.pstb_decode_standardize(store.tensor.data_ptr(), store.ld, n_iid, n_sid, full[0], full[1], 0, _lib.STD_UNIT,
                                               float("nan"), float("nan"), 0, stats.data_ptr(), out.data_ptr(), _lib.F32, _lib.ORDER_F, stream))

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.wait_started()
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    launches0 = lib.pstb_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for a, b in ev:
        a.record()
        step()
        b.record()
    e1.record()
    barrier()
    t1 = time.perf_counter()
    launches = lib.pstb_launch_count() - launches0
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    per_launch = [a.elapsed_time(b) for a, b in ev]
    kern_ms = float(np.mean(per_launch))
    clocks = sampler.window(t0, t1) if sampler else None
    ms_per_step = total_ms / args.steps
    value = world * n_iid * n_sid / (ms_per_step * 1e-3)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    algo_bytes = n_sid * rec + 4 * n_iid * n_sid + 16 * n_sid
    achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
    traffic = args.ncu_traffic
    if traffic is None:
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))["k_read_f_cfg2_dram_bytes_per_launch"]
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "traffic_source": "ncu --set full capture of this kernel on this workload (profiles/r2_read_f_full.txt), bytes per launch", "kernel": "k_read_f<float,warp> (fused decode+stats+standardize)",
                "algorithmic_bytes_per_launch": algo_bytes, "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s",
                "peak_note": "the measured peak is a torch copy (reads = writes); this kernel writes 94 % of its bytes, and a fraction slightly above 1 is "
                             "physical: ncu reads 6.58-6.67 TB/s = 80-82 % of its own DRAM peak with sm__cycles_active at 99.8 % of elapsed (profiles/r2_read_f_full.txt)"}

    # ---- spot parity of the timed configuration against the oracle (not timed) ----
    lib_o = _oracle_lib()
    sample_sid = min(args.cpu_sample_sid, n_sid)
    packed_sample = store.tensor[:sample_sid, :rec].contiguous().cpu().numpy()
    parity = None
    e2e = None
    cpu_baseline = None
    if rank == 0:
        threads = os.cpu_count() or 1
        cpu_s, cpu_out, cpu_stats = cpu_decode_standardize(lib_o, packed_sample, n_iid, threads, reps=2)
        cpu_baseline = {"value": n_iid * sample_sid / cpu_s, "unit": "genotypes/s", "cores": threads, "kind": "port",
                        "sample": "first {0} SNPs x {1} iids of the same synthetic store; oracle/c/pst_oracle.c decode+standardize (OpenMP)".format(sample_sid, n_iid)}
        chk = min(sample_sid, 2048)
        got = out[:chk].t().cpu().numpy()
        parity = {"checked_snps": chk, "max_abs_diff_vs_oracle": float(np.max(np.abs(got - cpu_out[:, :chk]))),
                  "stats_max_rel_diff": float(np.max(np.abs(stats[:chk].cpu().numpy() - cpu_stats[:chk]) / np.maximum(1e-300, np.abs(cpu_stats[:chk]))))}
        del cpu_out

    # ---- end to end through the host-buffer C ABI (every rank, pinned host buffers) ----
    if args.e2e:
        e2e = run_e2e(args, torch, lib, _lib, store, stats, n_iid, n_sid, rec, world, rank, barrier, max_over_ranks)
    api_e2e = None
    if args.api_e2e and rank == 0 and world == 1:
        try:
            api_e2e = run_api_e2e(torch, store, n_iid, n_sid, rec)
        except Exception as e:                                      # e.g. no room for the 2.5 GB file: the leg is reported as missing, the line survives
            api_e2e = {"unavailable": "{0}: {1}".format(type(e).__name__, e)}
    # ---- the other BASELINE configurations (single-GPU legs run at N = 1 only; cfg5 needs the memory of 8 GPUs) ----
    c_order = gather = kernel_missing = cfg5 = None
    if args.extra_legs and world == 1:
        c_order = run_c_order_leg(args, torch, dev, _lib, peaks, store, out)
    del out, store, stats
    torch.cuda.empty_cache()
    if args.extra_legs and world == 1:
        gather = run_gather_leg(args, torch, dev, _lib, peaks)
    # ---- SnpKernel (cfg3) ----
    kernel = None
    if args.kernel:
        kernel = run_kernel_workload(args, torch, dist, dev, _lib, rank, world, barrier, max_over_ranks, peaks, sampler)
        if args.extra_legs:
            # the same kernel on data WITH missing genotypes (5 %, like cfg4) and Beta(1,25): round 1 fell back to the 3-term split here
            kernel_missing = run_kernel_workload(args, torch, dist, dev, _lib, rank, world, barrier, max_over_ranks, peaks, sampler, n=args.kernel_n,
                                                 m=args.kernel_missing_m, missing=0.05, spec=("unit",), label="cfg3 shape with missing data", with_e2e=False, with_cpu=False)
            kernel_missing["beta"] = run_kernel_workload(args, torch, dist, dev, _lib, rank, world, barrier, max_over_ranks, peaks, None, n=args.kernel_n,
                                                         m=args.kernel_missing_m, missing=0.05, spec=("beta", 1, 25), label="cfg3 shape with missing data", with_e2e=False, with_cpu=False)
    if args.extra_legs and world >= 8:
        cfg5 = run_cfg5(args, torch, dist, dev, _lib, rank, world, barrier, max_over_ranks, peaks)
    if sampler:
        sampler.close()

    if rank == 0:
        line = {
            "metric": "genotypes/s decoded+standardized", "value": value, "unit": "genotypes/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg2: synthetic .bed 10000 iids x 1000000 SNPs per GPU, decode + Unit standardize float32, F order",
                       "packed_bytes": n_sid * rec, "out_bytes": 4 * n_iid * n_sid, "l2": "inputs (2.5 GB) and outputs (40 GB) larger than L2; no flush needed",
                       "sharding": "SNP ranges, one cfg2-sized shard per GPU, no collective",
                       "numa": (None if world == 1 else ("rank 0 bound to the CPUs / memory of its GPU's NUMA node {0} (every rank to its own)".format(NUMA_NODE)
                                                         if NUMA_NODE is not None and NUMA_NODE >= 0 else "no NUMA binding (none exposed by the host, or --no-numa-bind)"))},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "kernel_ms_per_launch": kern_ms, "kernel_ms_best": float(np.min(per_launch)), "kernel_ms_median": float(np.median(per_launch)),
            "parity_spot_check": parity,
        }
        if kernel is not None:
            line["kernel"] = kernel
        if api_e2e is not None:
            line["e2e_python_api"] = api_e2e
        for key, val in (("kernel_missing", kernel_missing), ("gather", gather), ("c_order", c_order), ("cfg5", cfg5)):
            if val is not None:
                line[key] = val
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, torch, lib, _lib, store, stats, n_iid, n_sid, rec, world, rank, barrier, max_over_ranks):
    """cfg2 through the host-buffer C ABI.  One GPU: the whole 10 000 x 1 000 000 workload (2.5 GB in, 40 GB out, pinned).
    N GPUs: the same single workload split by SNP range over the ranks (rank r streams SNPs [r*M/N, (r+1)*M/N) of its
    shard), so the host holds one 40 GB result in total instead of N of them."""
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    lo, hi = rank * n_sid // world, (rank + 1) * n_sid // world
    m = hi - lo
    shrink = 1
    while True:                                        # a box short of lockable memory gets a stated prefix instead of a crash
        h_out_p = lib.pstb_host_alloc(n_iid * m * 4)
        h_pk_p = lib.pstb_host_alloc(m * rec) if h_out_p else None
        if h_out_p and h_pk_p:
            break
        if h_out_p:
            lib.pstb_host_free(h_out_p)
        if m < 20000:
            raise RuntimeError("pinned host allocation failed: " + _lib.last_error())
        m //= 4
        shrink *= 4
    hi = lo + m
    h_packed = np.ctypeslib.as_array(ctypes.cast(h_pk_p, ctypes.POINTER(ctypes.c_uint8)), shape=(m, rec))
    step_rows = max(1, (1 << 28) // rec)
    for s0 in range(0, m, step_rows):
        h_packed[s0:s0 + step_rows] = store.tensor[lo + s0:lo + min(m, s0 + step_rows), :rec].cpu().numpy()
    h_stats = np.empty((m, 2), dtype=np.float64)

    def e2e_step():
        _lib.check(lib.pstb_read_host(h_pk_p, n_iid, m, None, n_iid, None, m, 0, _lib.STD_UNIT, float("nan"), float("nan"), 0,
                                      h_stats.ctypes.data, h_out_p, _lib.F32, _lib.ORDER_F))

    e2e_step()                                                                      # warm-up (allocates the chunk ring)
    barrier()
    t0e = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    dt_e = max_over_ranks((time.perf_counter() - t0e) / e2e_steps)
    h_out = np.ctypeslib.as_array(ctypes.cast(h_out_p, ctypes.POINTER(ctypes.c_float)), shape=(m, n_iid))
    e2e_ok = bool(np.array_equal(h_stats[:256], stats[lo:lo + 256].cpu().numpy())) and bool(np.isfinite(h_out[-1]).all())
    done_sid = int(max_over_ranks(float(m))) * world if shrink > 1 else n_sid        # SNPs actually streamed per step over all ranks
    res = {"value": n_iid * done_sid / dt_e, "unit": "genotypes/s", "h2d_bytes_per_step": done_sid * rec, "d2h_bytes_per_step": n_iid * done_sid * 4 + 16 * done_sid,
           "ms_per_step": dt_e * 1e3, "steps": e2e_steps, "api": "pstb_read_host (host-buffer C ABI), pinned host buffers, 64 MiB chunks on 4 streams",
           "work": ("one cfg2 workload in total" if shrink == 1 else "the first 1/{0} of the SNPs of one cfg2 workload (pinned host memory ran short)".format(shrink))
                   + ("" if world == 1 else ", SNP ranges split over the {0} ranks; bytes are totals over ranks".format(world)),
           "stats_match_device_run": e2e_ok}
    del h_out, h_packed
    lib.pstb_host_free(h_out_p)
    lib.pstb_host_free(h_pk_p)
    lib.pstb_host_release()
    return res


def run_api_e2e(torch, store, n_iid, n_sid, rec):
    """The user-facing call: a .bed file on disk -> Bed(...).read(dtype=float32, standardizer=Unit()) -> pageable NumPy array."""
    import tempfile
    from pysnptools_b200 import Bed, Unit
    d = tempfile.mkdtemp(prefix="pstb_bench_")
    path = os.path.join(d, "cfg2.bed")
    with open(path, "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01]))
        step_rows = max(1, (1 << 28) // rec)
        for s0 in range(0, n_sid, step_rows):
            f.write(store.tensor[s0:s0 + step_rows, :rec].contiguous().cpu().numpy().tobytes())
    iid = np.array([["f", str(k)] for k in range(n_iid)])
    sid = np.arange(n_sid).astype(str)
    pos = np.zeros((n_sid, 3))
    bed = Bed(path, count_A1=False, iid=iid, sid=sid, pos=pos)                 # labels given: no .fam/.bim parsing (bed.py:127-135)
    t = []
    for _ in range(2):
        t0 = time.perf_counter()
        data = bed.read(order="F", dtype=np.float32, standardizer=Unit())
        t.append(time.perf_counter() - t0)
        ok = bool(np.isfinite(data.val[:, -1]).all())
        del data
    from pysnptools_b200.util import pinned_empty
    buf = pinned_empty((n_iid, n_sid), dtype=np.float32, order="F")
    tp = []
    for _ in range(2):
        t0 = time.perf_counter()
        bed.read(order="F", dtype=np.float32, standardizer=Unit(), out=buf)
        tp.append(time.perf_counter() - t0)
    del buf
    os.remove(path)
    os.rmdir(d)
    return {"value": n_iid * n_sid / min(t), "unit": "genotypes/s", "seconds": t, "finite": ok,
            "api": "Bed(file).read(order='F', dtype=float32, standardizer=Unit()) -> fresh pageable NumPy array (file in the page cache)",
            "into_pinned_out": {"value": n_iid * n_sid / min(tp), "seconds": tp, "api": "the same call with out=pinned_empty(...)"}}


def run_kernel_api_e2e(torch, store, K_dev, n, m, rec):
    """cfg3 through the user-facing call: a .bed file on disk -> SnpKernel(Bed(file), Unit()).read(dtype=float32) -> KernelData (NumPy).
    This is the call patch_reference() binds into the reference (pstb_snp_kernel_host on the memory-mapped file: pageable in, pageable out)."""
    import tempfile
    from pysnptools_b200 import Bed, SnpKernel, Unit
    d = tempfile.mkdtemp(prefix="pstb_bench_")
    path = os.path.join(d, "cfg3.bed")
    with open(path, "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01]))
        step_rows = max(1, (1 << 28) // rec)
        for s0 in range(0, m, step_rows):
            f.write(store.tensor[s0:s0 + step_rows, :rec].contiguous().cpu().numpy().tobytes())
    iid = np.array([["f", str(k)] for k in range(n)])
    bed = Bed(path, count_A1=False, iid=iid, sid=np.arange(m).astype(str), pos=np.zeros((m, 3)))
    t = []
    worst = None
    for _ in range(2):
        t0 = time.perf_counter()
        kd = SnpKernel(bed, Unit()).read(dtype=np.float32)
        t.append(time.perf_counter() - t0)
        if worst is None:
            T = (n + 255) // 256
            worst = 0.0
            for I, J in ((0, 0), (T - 1, T - 1), (T // 2, T // 3), (T - 1, 0)):
                a = kd.val[I * 256:I * 256 + 256, J * 256:J * 256 + 256].astype(np.float64)
                b = K_dev[I * 256:I * 256 + 256, J * 256:J * 256 + 256].double().cpu().numpy()
                worst = max(worst, float(np.linalg.norm(a - b) / max(1e-300, np.linalg.norm(b))))
        del kd
    os.remove(path)
    os.rmdir(d)
    return {"value": 2.0 * n * n * m / min(t) / 1e12, "unit": "TFLOP/s", "seconds": t,
            "api": "SnpKernel(Bed(file), Unit()).read(dtype=float32) -> KernelData in pageable NumPy memory (file in the page cache); the call patch_reference() binds",
            "worst_rel_frobenius_vs_device_K_on_4_blocks": worst}


def pick_blocks(n, count, seed):
    """`count` distinct 256-row blocks: the first, the last (ragged) one and a seeded random choice of the others."""
    T = (n + 255) // 256
    count = min(count, T)
    rng = np.random.default_rng(seed)
    chosen = {0, T - 1}
    for b in rng.permutation(T):
        if len(chosen) >= count:
            break
        chosen.add(int(b))
    return sorted(chosen)


def run_kernel_workload(args, torch, dist, dev, _lib, rank, world, barrier, max_over_ranks, peaks, sampler=None, n=None, m=None,
                        missing=0.0, spec=("unit",), label="cfg3", with_e2e=True, with_cpu=True):
    """SnpKernel on n x m (default cfg3: 50 000 x 500 000, Unit), SNP-sharded over the ranks, one NCCL all-reduce of K; the result is
    compared with the CPU oracle on sampled 256 x 256 tiles (after the all-reduce at N > 1)."""
    n, m = (n or args.kernel_n, m or args.kernel_m)
    m_lo, m_hi = rank * m // world, (rank + 1) * m // world
    store = gen_store_device(dev, torch, n, m_hi - m_lo, seed=2000 + rank + (7000 if missing else 0), missing_rate=missing)
    K = torch.zeros((n, n), dtype=torch.float32, device="cuda")
    chunk = args.kernel_chunk
    # N GPUs: the partial kernels live in compact lower-triangular tile storage, so the all-reduce moves the triangle only
    compact = world > 1 and not args.square_allreduce
    tiles = torch.zeros((len(dev.kernel_tile_coords(n)), 256, 256), dtype=torch.float32, device="cuda") if compact else None

    marks = []
    low_term = dev.low_term_for(m, n, spec)                       # SNP shards: the low-term mode follows the whole kernel's SNP count
    stats_box = [None]

    def step():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        if compact and not args.serial_allreduce:
            # the NCCL all-reduce of finished tile bands runs on a side stream under the last SNP chunk's multiplication
            _k, stats_box[0] = parallel.snp_kernel_sharded_overlapped(store, n, m, None, spec, chunk=chunk, tiles=tiles, K=K,
                                                                      bands=args.allreduce_bands, reserve_sms=args.allreduce_sms, tail_chunks=args.allreduce_tail_chunks)
            for e in ev[1:]:
                e.record()
            marks.append(ev)
            return
        u_box = [None]
        if compact:
            # the rank-one vector of the exact-dosage path is deferred: all-reduced with the tiles (n doubles) and added during the expansion
            _t, _c, stats_box[0], u_box[0] = dev.snp_kernel_tiles(store, chunk=chunk, tiles=tiles, accumulate=False, low_term=low_term, standardizer=spec,
                                                                 defer_rank1=True)
        else:
            _k, stats_box[0] = dev.snp_kernel(store, K=K, accumulate=False, chunk=chunk, mirror=(world == 1), low_term=low_term, standardizer=spec)
        ev[1].record()
        if compact:
            # ONE reduction of the compact triangle over NVLink, issued in slices so that the expansion of a reduced slice into the
            # square K runs while the next slice is still being reduced
            parallel.allreduce_tiles_and_expand(tiles, n, K, slices=args.allreduce_slices, u=u_box[0])
            ev[2].record()
        elif world > 1:
            dist.all_reduce(K)
            ev[2].record()
            _lib.check(_lib.lib.pstb_mirror_lower(K.data_ptr(), n, n, torch.cuda.current_stream().cuda_stream))
        else:
            ev[2].record()
        ev[3].record()
        marks.append(ev)

    step()
    barrier()
    marks.clear()
    l0 = _lib.lib.pstb_launch_count()
    steps = max(1, min(args.steps, args.kernel_steps))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tk0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    tk1 = time.perf_counter()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    kclocks = sampler.window(tk0, tk1) if sampler else None
    launches = int(_lib.lib.pstb_launch_count() - l0)
    breakdown = {"compute_ms": float(np.mean([e[0].elapsed_time(e[1]) for e in marks])),
                 "allreduce_ms_incl_wait_for_slowest_rank": float(np.mean([e[1].elapsed_time(e[2]) for e in marks])),
                 "mirror_ms": float(np.mean([e[2].elapsed_time(e[3]) for e in marks])),
                 "allreduce": (("compact lower-triangular tiles ({0:.2f} GB) in {1} bands, all-reduced + expanded on a side stream while the last SNP chunks are multiplied "
                                "({2} SMs left to NCCL; the last {3} chunks are multiplied band-major); compute_ms is the whole overlapped step").format(tiles.numel() * 4 / 1e9, args.allreduce_bands, args.allreduce_sms, args.allreduce_tail_chunks)
                               if (compact and not args.serial_allreduce) else
                               "compact lower-triangular tiles ({0:.2f} GB) in {1} slices, each expanded to the square matrix while the next is reduced (expansion of the last one in the 'mirror' slot)".format(tiles.numel() * 4 / 1e9, args.allreduce_slices)
                               if compact else ("square matrix" if world > 1 else "none"))}
    tflops = 2.0 * n * n * m / (ms * 1e-3) / 1e12
    diag = float(K.diagonal().double().mean().item())
    # ---- parity of THIS result (every rank holds the all-reduced K) against the CPU oracle on sampled tiles ----
    parity = None
    if args.kernel_parity_blocks > 0:
        blocks = pick_blocks(n, args.kernel_parity_blocks, seed=n + m)

        def fetch(I, J):
            if rank != 0:
                return np.zeros((256, 256))
            sub = K[I * 256:I * 256 + 256, J * 256:J * 256 + 256].double().cpu().numpy()
            out = np.zeros((256, 256))
            out[: sub.shape[0], : sub.shape[1]] = sub
            return out
        parity = sampled_tile_parity(torch, dist, world, rank, _oracle_lib(), store, stats_box[0], n, spec, fetch, blocks)
        parity["symmetric"] = bool(torch.equal(K[:512, -512:], K[-512:, :512].t()))
    e2e = run_kernel_e2e(args, torch, dist, dev, _lib, store, K, tiles, n, m, m_hi - m_lo, chunk, rank, world, barrier, max_over_ranks, low_term) if (args.e2e and with_e2e) else None
    api_e2e = None
    if args.e2e and with_e2e and args.api_e2e and world == 1 and spec == ("unit",):
        try:
            api_e2e = run_kernel_api_e2e(torch, store, K, n, m, (n + 3) // 4)
        except Exception as e:                                      # e.g. no room for the 6.25 GB file
            api_e2e = {"unavailable": "{0}: {1}".format(type(e).__name__, e)}
        _lib.lib.pstb_host_release()
    cpu_baseline = None
    if rank == 0 and world == 1 and args.kernel_cpu and with_cpu:
        del K
        torch.cuda.empty_cache()
        cpu_baseline = cpu_kernel_baseline(_oracle_lib(), n, args.ref_kernel_sid, os.cpu_count() or 1)
        K = torch.zeros((1, 1), device="cuda")
    fp8lo = low_term == "fp8"
    t256 = (n + 255) // 256
    ntiles = t256 * (t256 + 1) // 2                                                # lower-triangular 256 x 256 tiles per rank
    # every chunk takes the 2-term exact-dosage GEMM, with or without missing genotypes (3 terms with PSTB_SYRK_3TERM=1)
    terms = 3 if os.environ.get("PSTB_SYRK_3TERM", "0") not in ("", "0") or os.environ.get("PSTB_SYRK_V1", "0") not in ("", "0") else 2
    # tensor-pipe work in fp16-equivalent terms: an fp8 (e4m3) MMA term costs half the cycles of an fp16 one
    pipe_terms = 1.5 if (terms == 2 and fp8lo) else float(terms)
    executed = pipe_terms * 2.0 * 256 * 256 * ntiles * (((m_hi - m_lo) + 63) // 64 * 64) / (ms * 1e-3) / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    burst = float(peaks.get("bf16_tflops", 1650.0))
    std_name = "Unit" if spec[0] == "unit" else "Beta({0},{1})".format(spec[1], spec[2])
    res = {"metric": "SnpKernel TFLOP/s (2*N^2*M)", "value": tflops, "e2e": e2e, "cpu_baseline": cpu_baseline, "unit": "TFLOP/s", "n_gpus": world, "ms_per_step": ms, "steps": steps,
           "scaling": "strong", "higher_is_better": True,
           "config": {"workload": "{0}: synthetic .bed {1} iids x {2} SNPs, {3:.0f} % missing, SnpKernel({4}), K fp32, SNP-sharded + NCCL allreduce".format(label, n, m, 100 * missing, std_name),
                      "chunk_snps": chunk or dev.default_kernel_chunk(n, m_hi - m_lo), "split": ("exact fp16 left plane (g - mu', missing -> mu - mu') x fp16 high part of the weighted centred value + e4m3 x e4m3 low term on the fp8 pipe (1 fp16 + 1 fp8 MMA term per k-step = 1.5 fp16-equivalent terms)"
                                if (terms == 2 and fp8lo) else "exact fp16 left plane x fp16 hi/lo right plane ({0} MMA terms per k-step)".format(terms))
                               + "; lower-triangular 256x256 tiles on CTA pairs (tcgen05 cta_group::2); K leaves through TMA bulk tensor store / reduce-add",
                      "low_term": "fp8" if fp8lo else "fp16"},
           "roofline": {"bound": "tensor", "achieved": executed, "peak": peak, "unit": "TFLOP/s", "frac": executed / peak, "frac_of_burst_peak": executed / burst,
                        "note": "executed tensor-pipe work per rank in fp16-equivalent flops ({0} terms x lower-triangular tiles; an fp8 term counts half) / time; peak = MEASURED_PEAKS bf16_tflops_sustained (a step lasts seconds under the power cap); burst peak {1:.0f}".format(pipe_terms, burst)},
           "gpu_launches": launches, "mean_diag_over_M": diag / m, "parity": parity, "rank0_breakdown": breakdown, "clocks": kclocks}
    if api_e2e is not None:
        res["e2e_python_api"] = api_e2e
    del store, K, tiles
    torch.cuda.empty_cache()
    return res


def run_kernel_e2e(args, torch, dist, dev, _lib, store, K, tiles, n, m, m_local, chunk, rank, world, barrier, max_over_ranks, low_term="default"):
    """cfg3 end to end from HOST buffers: pinned packed bytes -> K in pinned host memory, copies inside the timed region.
    One GPU: ONE call of pstb_snp_kernel_host (the C ABI a bed_reader-style binding would use).  N GPUs: every rank uploads its
    SNP shard, computes its partial K, NCCL all-reduce, rank 0 copies the float32 K to the host."""
    import ctypes
    lib = _lib.lib
    rec = (n + 3) // 4
    chunk = chunk or dev.default_kernel_chunk(n, m_local)
    h_pk = lib.pstb_host_alloc(m_local * rec)
    h_K = lib.pstb_host_alloc(n * n * 4) if rank == 0 else None
    if not h_pk or (rank == 0 and not h_K):
        return {"unavailable": "pinned host allocation failed: " + _lib.last_error()}
    h_packed = np.ctypeslib.as_array(ctypes.cast(h_pk, ctypes.POINTER(ctypes.c_uint8)), shape=(m_local, rec))
    step_rows = max(1, (1 << 28) // rec)
    for s0 in range(0, m_local, step_rows):
        h_packed[s0:s0 + step_rows] = store.tensor[s0:min(m_local, s0 + step_rows), :rec].cpu().numpy()
    h_stats = np.empty((m_local, 2), dtype=np.float64)
    stream = torch.cuda.current_stream().cuda_stream

    if world == 1:
        def step():
            _lib.check(lib.pstb_snp_kernel_host(h_pk, n, m_local, None, n, None, m_local, 0, _lib.STD_UNIT, float("nan"), float("nan"), 0,
                                                h_stats.ctypes.data, h_K, _lib.F32, chunk, dev._LOW_TERM[low_term]))
        api = "pstb_snp_kernel_host (host-buffer C ABI): pinned packed bytes in, pinned float32 K out (packed slices cross PCIe under the SYRK, finished row ranges of K under its last chunks)"
    else:
        t_pk = torch.from_numpy(h_packed)
        d_tight = torch.empty((m_local, rec), dtype=torch.uint8, device="cuda")
        # ONE host copy of K shared by the ranks (a file mapping; every rank page-locks and fills its own row band): the finished kernel
        # leaves over all the PCIe links instead of rank 0's alone
        shared = None
        r0, r1 = rank * n // world, (rank + 1) * n // world
        try:
            base = "/dev/shm" if os.path.isdir("/dev/shm") and os.statvfs("/dev/shm").f_bavail * os.statvfs("/dev/shm").f_frsize > n * n * 4 + (1 << 30) else None
            ok_all = max_over_ranks(0.0 if base else 1.0) == 0.0
            if ok_all:
                path = os.path.join(base, "pstb_bench_K_{0}.bin".format(os.environ.get("MASTER_PORT", "0")))
                if rank == 0:
                    with open(path, "wb") as f:
                        f.truncate(n * n * 4)
                barrier()
                mm = np.memmap(path, dtype=np.float32, mode="r+", shape=(n, n))
                band = mm[r0:r1]
                rcr = torch.cuda.cudart().cudaHostRegister(band.ctypes.data, band.nbytes, 0)
                if max_over_ranks(0.0 if int(rcr) == 0 else 1.0) == 0.0:
                    shared = (path, mm, band, torch.from_numpy(band))
                elif int(rcr) == 0:
                    torch.cuda.cudart().cudaHostUnregister(band.ctypes.data)
        except Exception:
            shared = None
        if max_over_ranks(0.0 if shared is not None else 1.0) != 0.0:
            shared = None
        t_K = torch.from_numpy(np.ctypeslib.as_array(ctypes.cast(h_K, ctypes.POINTER(ctypes.c_float)), shape=(n, n))) if rank == 0 else None

        def step():
            d_tight.copy_(t_pk, non_blocking=True)                                  # H2D of this rank's SNP shard
            store.tensor[:, :rec].copy_(d_tight)                                    # re-pitch to the 16-byte record stride
            if tiles is not None and not args.serial_allreduce:
                parallel.snp_kernel_sharded_overlapped(store, n, m, None, ("unit",), chunk=chunk, tiles=tiles, K=K, bands=args.allreduce_bands,
                                                       reserve_sms=args.allreduce_sms, tail_chunks=args.allreduce_tail_chunks)
            elif tiles is not None:
                _t, _c, _s, u_e = dev.snp_kernel_tiles(store, chunk=chunk, tiles=tiles, accumulate=False, low_term=low_term, defer_rank1=True)
                parallel.allreduce_tiles_and_expand(tiles, n, K, slices=args.allreduce_slices, u=u_e)
            else:
                dev.snp_kernel(store, K=K, accumulate=False, chunk=chunk, mirror=False, low_term=low_term)
                dist.all_reduce(K)
                _lib.check(lib.pstb_mirror_lower(K.data_ptr(), n, n, stream))
            if shared is not None:
                shared[3].copy_(K[r0:r1], non_blocking=True)                        # D2H of this rank's row band of the finished kernel
            elif rank == 0:
                t_K.copy_(K, non_blocking=True)                                     # D2H of the finished kernel
            torch.cuda.synchronize()
        api = ("per rank: pinned packed shard -> HBM, SNP-sharded SnpKernel with the NCCL all-reduce of the compact triangle overlapped, every rank copies its row band of the float32 K "
               "into ONE shared page-locked host matrix" if shared is not None else
               "per rank: pinned packed shard -> HBM, pstb_snp_kernel, NCCL all-reduce, rank 0 copies float32 K to pinned host memory")
    step()
    barrier()
    steps = max(1, min(args.steps, args.kernel_steps))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    dt = max_over_ranks((time.perf_counter() - t0) / steps)
    ok = None
    if world > 1:
        barrier()
    if rank == 0:
        hk = shared[1] if (world > 1 and shared is not None) else np.ctypeslib.as_array(ctypes.cast(h_K, ctypes.POINTER(ctypes.c_float)), shape=(n, n))
        # the host copy against the device-resident K of the timed leg (itself checked against the oracle on sampled tiles):
        # diagonal and off-diagonal 256 x 256 blocks to 2e-6 relative Frobenius (slice boundaries may regroup the fp32 sums), symmetry
        T = (n + 255) // 256
        worst = 0.0
        for I, J in ((0, 0), (T - 1, T - 1), (T // 2, T // 3), (T - 1, 0), (T // 2, T // 2)):
            a = hk[I * 256:I * 256 + 256, J * 256:J * 256 + 256].astype(np.float64)
            b = K[I * 256:I * 256 + 256, J * 256:J * 256 + 256].double().cpu().numpy()
            worst = max(worst, float(np.linalg.norm(a - b) / max(1e-300, np.linalg.norm(b))))
        ok = {"worst_rel_frobenius_vs_device_K_on_5_blocks": worst, "within_2e-6": bool(worst <= 2e-6),
              "symmetric_block": bool(np.array_equal(hk[:256, -256:], hk[-256:, :256].T))}
        del hk
        lib.pstb_host_free(h_K)
    if world > 1 and shared is not None:
        barrier()
        torch.cuda.cudart().cudaHostUnregister(shared[2].ctypes.data)
        path = shared[0]
        shared = None
        barrier()
        if rank == 0:
            try:
                os.remove(path)
            except OSError:
                pass
    del h_packed
    lib.pstb_host_free(h_pk)
    lib.pstb_host_release()
    return {"value": 2.0 * n * n * m / dt / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": m * rec, "d2h_bytes_per_step": n * n * 4 + (16 * m if world == 1 else 0),
            "ms_per_step": dt * 1e3, "steps": steps, "api": api, "result_check": ok}


def run_cfg5(args, torch, dist, dev, _lib, rank, world, barrier, max_over_ranks, peaks):
    """cfg5: streamed decode + standardize + K for N = 500 000 (K = 1 TB): K-tile sharding, packed store replicated, no collective."""
    n, m = args.cfg5_n, args.cfg5_m
    coords = dev.kernel_tile_coords(n, rank, world)
    need = len(coords) * 256 * 256 * 4
    free, _total = torch.cuda.mem_get_info()
    rec = (n + 3) // 4
    chunk = 4096
    budget = need + m * (rec + 16) + int(_lib.lib.pstb_kernel_workspace_bytes(n, chunk)) + (2 << 30)
    if budget > free:
        return {"metric": "SnpKernel TFLOP/s (2*N^2*M)", "unavailable": "cfg5 needs {0:.0f} GB per GPU at {1} GPUs ({2:.0f} GB free); run with more GPUs or --cfg5-n".format(budget / 1e9, world, free / 1e9)}
    store = gen_store_device(dev, torch, n, m, seed=5000)                       # every rank holds the whole packed store (12.5 GB)
    tiles = torch.zeros((len(coords), 256, 256), dtype=torch.float32, device="cuda")

    def step():
        return dev.snp_kernel_tiles(store, rank=rank, world=world, chunk=chunk, tiles=tiles, accumulate=False)

    _, _, stats = step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    # parity on sampled tiles against the CPU oracle: >= 64 tiles in total (all pairs of `cfg5_blocks` row blocks), each compared on the
    # rank that owns it; the float64 reference is built from SNP shards of the (replicated) store, one shard per rank, and summed
    blocks = pick_blocks(n, args.cfg5_blocks, seed=n)
    index_of = {(int(c[0]), int(c[1])): t for t, c in enumerate(coords)}

    def fetch(I, J):
        t = index_of.get((I, J))
        return None if t is None else tiles[t].double().cpu().numpy()
    lo, hi = rank * m // world, (rank + 1) * m // world
    shard = dev.PackedStore(store.tensor[lo:hi], n, hi - lo)
    parity = sampled_tile_parity(torch, dist, world, rank, _oracle_lib(), shard, stats[lo:hi], n, ("unit",), fetch, blocks, owner_of=True)
    low_term = dev.low_term_for(m, n)
    t256 = (n + 255) // 256
    executed = (1.5 if low_term == "fp8" else 2.0) * 2.0 * 256 * 256 * (t256 * (t256 + 1) // 2) * ((m + 63) // 64 * 64) / (ms * 1e-3) / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0)) * world
    res = {"metric": "SnpKernel TFLOP/s (2*N^2*M)", "value": 2.0 * n * n * m / (ms * 1e-3) / 1e12, "unit": "TFLOP/s", "n_gpus": world, "ms_per_step": ms,
           "config": {"workload": "cfg5: synthetic .bed {0} iids x {1} SNPs, streamed decode+standardize+K, K-tile sharded over {2} GPUs (no collective)".format(n, m, world),
                      "tiles_per_rank": int(len(coords)), "tile_bytes_per_rank": need, "chunk_snps": chunk, "low_term": low_term},
           "roofline": {"bound": "tensor", "achieved": executed, "peak": peak, "unit": "TFLOP/s", "frac": executed / peak,
                        "note": "executed fp16-equivalent tensor work of all ranks / time; peak = ranks x MEASURED_PEAKS bf16_tflops_sustained"},
           "parity": parity}
    del store, tiles, shard
    torch.cuda.empty_cache()
    return res


def run_gather_leg(args, torch, dev, _lib, peaks):
    """cfg4 (BASELINE configs[3]): 100 000 iids x 200 000 SNPs, 5 % missing, Beta(1,25), random unsorted iid / sid subsets
    (rng = default_rng(1); permutation(N)[:N//2], permutation(M)[:M//2] -- SURVEY 8d), float32, F order.  One fused launch."""
    lib = _lib.lib
    n, m = args.cfg4_n, args.cfg4_m
    rng = np.random.default_rng(1)
    iid_idx, sid_idx = rng.permutation(n)[: n // 2], rng.permutation(m)[: m // 2]
    store = gen_store_device(dev, torch, n, m, seed=4000, missing_rate=0.05)
    I, S = dev.Selection(iid_idx, n, "cuda"), dev.Selection(sid_idx, m, "cuda")
    out = torch.empty((S.n, I.n), dtype=torch.float32, device="cuda")
    stats = torch.empty((S.n, 2), dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def step():
        _lib.check(lib.pstb_decode_standardize(store.tensor.data_ptr(), store.ld, n, m, I.axis(), S.axis(), 0, _lib.STD_BETA, 1.0, 25.0, 0,
                                               stats.data_ptr(), out.data_ptr(), _lib.F32, _lib.ORDER_F, st))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    steps = max(3, min(args.steps, 10))
    ts = []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.mean(ts))
    rec = (n + 3) // 4
    algo = S.n * rec + 4 * I.n * S.n + 16 * S.n
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # spot parity: the first 1 024 selected SNPs against the oracle (bit pattern of NaN handling + values to 1e-6 relative)
    lib_o = _oracle_lib()
    chk = min(1024, S.n)
    threads = os.cpu_count() or 1
    packed = np.ascontiguousarray(store.tensor[torch.as_tensor(sid_idx[:chk], device="cuda"), :rec].cpu().numpy())
    ref = np.empty((I.n, chk), dtype=np.float64, order="F")
    rst = np.empty((chk, 2), dtype=np.float64)
    p = ctypes.c_void_p
    ii64 = np.ascontiguousarray(iid_idx, dtype=np.int64)
    lib_o.pst_oracle_decode_f64(p(packed.ctypes.data), ctypes.c_int64(rec), ctypes.c_int64(n), ctypes.c_int64(chk), p(ii64.ctypes.data), ctypes.c_int64(I.n),
                                None, ctypes.c_int64(chk), 0, 0, p(ref.ctypes.data), threads)
    lib_o.pst_oracle_standardize_f64(p(ref.ctypes.data), ctypes.c_int64(I.n), ctypes.c_int64(chk), 0, 1, ctypes.c_double(1.0), ctypes.c_double(25.0), 0,
                                     p(rst.ctypes.data), threads)
    got = out[:chk].t().double().cpu().numpy()
    scale = max(1e-300, float(np.max(np.abs(ref))))
    parity = {"checked_snps": chk, "max_abs_diff_over_max_abs": float(np.max(np.abs(got - ref)) / scale), "tolerance": 1e-6,
              "stats_max_rel_diff": float(np.max(np.abs(stats[:chk].cpu().numpy() - rst) / np.maximum(1e-300, np.abs(rst))))}
    res = {"metric": "genotypes/s decoded+standardized", "value": I.n * S.n / (ms * 1e-3), "unit": "genotypes/s", "ms_per_step": ms, "steps": steps,
           "config": {"workload": "cfg4: synthetic .bed {0} iids x {1} SNPs, 5 % missing, Beta(1,25), random unsorted {2} x {3} subset (default_rng(1) permutations), float32, F order".format(n, m, I.n, S.n)},
           "roofline": {"bound": "hbm", "achieved": algo / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": algo / (ms * 1e-3) / 1e9 / hbm_peak,
                        "algorithmic_bytes_per_launch": algo, "kernel": "k_read_f_gather4<float> (byte-interleaved gather of four records + masked-popcount statistics)"},
           "parity_spot_check": parity}
    del store, out, stats
    torch.cuda.empty_cache()
    return res


def run_c_order_leg(args, torch, dev, _lib, peaks, store, ref_out_f):
    """cfg2 with order='C' (sid fastest: the layout the reference's kernel loop asks for, snpreader.py:640): k_stats_dense + k_emit_c_wide."""
    lib = _lib.lib
    n, m = store.iid_count, args.c_order_m
    out = torch.empty((n, m), dtype=torch.float32, device="cuda")
    stats = torch.empty((m, 2), dtype=torch.float64, device="cuda")
    full = _lib.Axis(None, 0, 1, n), _lib.Axis(None, 0, 1, m)
    st = torch.cuda.current_stream().cuda_stream

    def step():
        _lib.check(lib.pstb_decode_standardize(store.tensor.data_ptr(), store.ld, n, store.sid_count, full[0], full[1], 0, _lib.STD_UNIT, float("nan"), float("nan"), 0,
                                               stats.data_ptr(), out.data_ptr(), _lib.F32, _lib.ORDER_C, st))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(max(3, min(args.steps, 10))):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.mean(ts))
    rec = (n + 3) // 4
    algo = 2 * m * rec + 4 * n * m + 16 * m                      # the packed records are read twice: statistics pass + emit pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    same = bool(torch.equal(out[:, :4096], ref_out_f[:4096].t())) if ref_out_f is not None else None
    res = {"metric": "genotypes/s decoded+standardized", "value": n * m / (ms * 1e-3), "unit": "genotypes/s", "ms_per_step": ms,
           "config": {"workload": "cfg2 shape in C order: {0} iids x {1} SNPs, decode + Unit, float32".format(n, m)},
           "roofline": {"bound": "hbm", "achieved": algo / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": algo / (ms * 1e-3) / 1e9 / hbm_peak,
                        "algorithmic_bytes_per_launch": algo, "kernel": "k_stats_dense + k_emit_c_wide<float>"},
           "bit_identical_to_F_order_result_on_4096_snps": same}
    del out, stats
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample-sid", type=int, default=100_000)
    ap.add_argument("--ref-sample-sid", type=int, default=50_000)
    ap.add_argument("--ref-kernel-sid", type=int, default=1024, help="SNPs in the CPU SnpKernel sample block (full N)")
    ap.add_argument("--no-kernel", dest="kernel", action="store_false")
    ap.add_argument("--cfg5", action="store_true", help="run only the cfg5 leg: K-tile sharded SnpKernel (500 000 x 100 000 across the ranks)")
    ap.add_argument("--cfg5-n", type=int, default=500_000)
    ap.add_argument("--cfg5-m", type=int, default=100_000)
    ap.add_argument("--cfg5-blocks", type=int, default=11, help="row blocks whose pairwise tiles (66 for 11) are compared with the CPU oracle")
    ap.add_argument("--cfg4-n", type=int, default=100_000)
    ap.add_argument("--cfg4-m", type=int, default=200_000)
    ap.add_argument("--c-order-m", type=int, default=CFG2["n_sid"])
    ap.add_argument("--kernel-missing-m", type=int, default=100_032, help="SNPs of the missing-data SnpKernel leg (cfg3 individuals)")
    ap.add_argument("--kernel-parity-blocks", type=int, default=11, help="row blocks whose pairwise 256x256 tiles (66 for 11) of K are compared with the CPU oracle; 0 = off")
    ap.add_argument("--no-extra-legs", dest="extra_legs", action="store_false", help="skip the cfg4 / C-order / missing-data / cfg5 legs")
    ap.add_argument("--no-api-e2e", dest="api_e2e", action="store_false", help="skip Bed(file).read(dtype=float32, standardizer=Unit()) -> NumPy through the Python layer")
    ap.add_argument("--only-kernel", action="store_true", help="experiments: run only the cfg3 SnpKernel leg and print its object")
    ap.add_argument("--no-numa-bind", dest="numa_bind", action="store_false", help="N > 1: do not bind each rank to the CPUs / memory of its GPU's NUMA node")
    ap.add_argument("--no-e2e", dest="e2e", action="store_false", help="profiling runs only: skip the host-buffer leg")
    ap.add_argument("--kernel-n", type=int, default=CFG3["n_iid"])
    ap.add_argument("--kernel-m", type=int, default=CFG3["n_sid"])
    ap.add_argument("--kernel-steps", type=int, default=2)
    ap.add_argument("--kernel-chunk", type=int, default=None)
    ap.add_argument("--square-allreduce", action="store_true", help="A/B: all-reduce the square K instead of the compact lower triangle")
    ap.add_argument("--serial-allreduce", action="store_true", help="A/B: all-reduce the compact triangle (in slices, expansion pipelined) after the whole multiplication instead of overlapping it with the last SNP chunks")
    ap.add_argument("--allreduce-slices", type=int, default=4, help="slices of the compact triangle's all-reduce; a reduced slice is expanded while the next is in flight")
    ap.add_argument("--allreduce-bands", type=int, default=8)
    ap.add_argument("--allreduce-sms", type=int, default=0, help="SMs the tail chunks' SYRK leaves idle for the overlapped NCCL all-reduce (0: the dynamic tile feed absorbs whatever the collective takes)")
    ap.add_argument("--allreduce-tail-chunks", type=int, default=4, help="SNP chunks multiplied band-major at the end, under which the all-reduce runs")
    ap.add_argument("--no-kernel-cpu", dest="kernel_cpu", action="store_false", help="skip the CPU SnpKernel sample (NumPy BLAS) of the kernel leg")
    ap.add_argument("--ncu-traffic", type=float, default=None, help="dram bytes per launch from profiles/ (ncu --set full), for the roofline object")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
