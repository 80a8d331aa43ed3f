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
    from pysnptools_b200 import _lib, device as dev
    lib = _lib.lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
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
        res = run_cfg5(args, torch, dist, dev, _lib, rank, world, barrier, max_over_ranks)
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
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))["k_read_f_cfg2_dram_bytes_per_launch"]
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "traffic_source": "ncu --set full capture of this kernel on this workload (profiles/r1_read_f_full.txt), bytes per launch", "kernel": "k_read_f<float,warp> (fused decode+stats+standardize)",
                "algorithmic_bytes_per_launch": algo_bytes, "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s"}

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
    if args.api_e2e and rank == 0:
        api_e2e = run_api_e2e(torch, store, n_iid, n_sid, rec)
    del out, store, stats
    torch.cuda.empty_cache()
    # ---- SnpKernel (cfg3) ----
    kernel = None
    if args.kernel:
        kernel = run_kernel_workload(args, torch, dist, dev, _lib, rank, world, barrier, max_over_ranks, peaks, sampler)
    if sampler:
        sampler.close()

    if rank == 0:
        line = {
            "metric": "genotypes/s decoded+standardized", "value": value, "unit": "genotypes/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg2: synthetic .bed 10000 iids x 1000000 SNPs per GPU, decode + Unit standardize float32, F order",
                       "packed_bytes": n_sid * rec, "out_bytes": 4 * n_iid * n_sid, "l2": "inputs (2.5 GB) and outputs (40 GB) larger than L2; no flush needed",
                       "sharding": "SNP ranges, one cfg2-sized shard per GPU, no collective"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "kernel_ms_per_launch": kern_ms, "kernel_ms_best": float(np.min(per_launch)), "kernel_ms_median": float(np.median(per_launch)),
            "parity_spot_check": parity,
        }
        if kernel is not None:
            line["kernel"] = kernel
        if api_e2e is not None:
            line["e2e_python_api"] = api_e2e
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


def run_kernel_workload(args, torch, dist, dev, _lib, rank, world, barrier, max_over_ranks, peaks, sampler=None):
    """cfg3: SnpKernel(Unit) on 50 000 x 500 000, SNP-sharded over the ranks, one NCCL all-reduce of K."""
    n, m = (args.kernel_n, args.kernel_m)
    m_lo, m_hi = rank * m // world, (rank + 1) * m // world
    store = gen_store_device(dev, torch, n, m_hi - m_lo, seed=2000 + rank)
    K = torch.zeros((n, n), dtype=torch.float32, device="cuda")
    chunk = args.kernel_chunk
    # N GPUs: the partial kernels live in compact lower-triangular tile storage, so the all-reduce moves the triangle only
    compact = world > 1 and not args.square_allreduce
    tiles = torch.zeros((len(dev.kernel_tile_coords(n)), 256, 256), dtype=torch.float32, device="cuda") if compact else None

    marks = []

    low_term = dev.low_term_for(m, n)                             # SNP shards: the low-term mode follows the whole kernel's SNP count

    def step():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        if compact:
            dev.snp_kernel_tiles(store, chunk=chunk, tiles=tiles, accumulate=False, low_term=low_term)
        else:
            dev.snp_kernel(store, K=K, accumulate=False, chunk=chunk, mirror=(world == 1), low_term=low_term)
        ev[1].record()
        if compact:
            dist.all_reduce(tiles)                                                   # sum of the partial triangles over NVLink
            ev[2].record()
            dev.kernel_from_tiles(tiles, n, K=K)                                     # -> full symmetric K
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
    breakdown = {"compute_ms": float(np.mean([e[0].elapsed_time(e[1]) for e in marks])),
                 "allreduce_ms_incl_wait_for_slowest_rank": float(np.mean([e[1].elapsed_time(e[2]) for e in marks])),
                 "mirror_ms": float(np.mean([e[2].elapsed_time(e[3]) for e in marks])),
                 "allreduce": ("compact lower-triangular tiles ({0:.2f} GB), expanded to the square matrix in the 'mirror' slot".format(tiles.numel() * 4 / 1e9)
                               if compact else ("square matrix" if world > 1 else "none"))}
    tflops = 2.0 * n * n * m / (ms * 1e-3) / 1e12
    diag = float(K.diagonal().double().mean().item())
    e2e = run_kernel_e2e(args, torch, dist, dev, _lib, store, K, tiles, n, m, m_hi - m_lo, chunk, rank, world, barrier, max_over_ranks, low_term) if args.e2e else None
    cpu_baseline = None
    if rank == 0 and world == 1 and args.kernel_cpu:
        del K
        torch.cuda.empty_cache()
        cpu_baseline = cpu_kernel_baseline(_oracle_lib(), n, args.ref_kernel_sid, os.cpu_count() or 1)
        K = torch.zeros((1, 1), device="cuda")
    fp8lo = low_term == "fp8"
    t256 = (n + 255) // 256
    tiles = t256 * (t256 + 1) // 2                                                 # lower-triangular 256 x 256 tiles per rank
    # synthetic cfg3 has no missing genotypes: every chunk takes the 2-term exact-dosage GEMM (3 terms with PSTB_SYRK_3TERM=1)
    terms = 3 if os.environ.get("PSTB_SYRK_3TERM", "0") not in ("", "0") or os.environ.get("PSTB_SYRK_V1", "0") not in ("", "0") else 2
    # tensor-pipe work in fp16-equivalent terms: an fp8 (e4m3) MMA term costs half the cycles of an fp16 one
    pipe_terms = 1.5 if (terms == 2 and fp8lo) else float(terms)
    executed = pipe_terms * 2.0 * 256 * 256 * tiles * (((m_hi - m_lo) + 63) // 64 * 64) / (ms * 1e-3) / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    return {"metric": "SnpKernel TFLOP/s (2*N^2*M)", "value": tflops, "e2e": e2e, "cpu_baseline": cpu_baseline, "unit": "TFLOP/s", "n_gpus": world, "ms_per_step": ms, "steps": steps,
            "config": {"workload": "cfg3: synthetic .bed {0} iids x {1} SNPs, SnpKernel(Unit), K fp32, SNP-sharded + NCCL allreduce".format(n, m),
                       "chunk_snps": chunk or dev.default_kernel_chunk(n, m_hi - m_lo), "split": ("exact fp16 dosage x fp16 high part of the weighted dosage + e4m3 x e4m3 low term on the fp8 pipe (1 fp16 + 1 fp8 MMA term per k-step = 1.5 fp16-equivalent terms)"
                                 if (terms == 2 and fp8lo) else "exact fp16 dosage x fp16 hi/lo weights ({0} MMA terms per k-step)".format(terms))
                                + "; 3-term hi/lo split when a chunk has missing data; lower-triangular 256x256 tiles on CTA pairs (tcgen05 cta_group::2)",
                       "low_term": "fp8" if fp8lo else "fp16"},
            "roofline": {"bound": "tensor", "achieved": executed, "peak": peak, "unit": "TFLOP/s", "frac": executed / peak,
                         "note": "executed tensor-pipe work per rank in fp16-equivalent flops ({0} terms x lower-triangular tiles; an fp8 term counts half) / time; peak = MEASURED_PEAKS bf16_tflops_sustained".format(pipe_terms)},
            "gpu_launches": int(_lib.lib.pstb_launch_count() - l0), "mean_diag_over_M": diag / m, "rank0_breakdown": breakdown, "clocks": kclocks}


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
        api = "pstb_snp_kernel_host (host-buffer C ABI): pinned packed bytes in, pinned float32 K out"
    else:
        t_pk = torch.from_numpy(h_packed)
        d_tight = torch.empty((m_local, rec), dtype=torch.uint8, device="cuda")
        t_K = torch.from_numpy(np.ctypeslib.as_array(ctypes.cast(h_K, ctypes.POINTER(ctypes.c_float)), shape=(n, n))) if rank == 0 else None

        def step():
            d_tight.copy_(t_pk, non_blocking=True)                                  # H2D of this rank's SNP shard
            store.tensor[:, :rec].copy_(d_tight)                                    # re-pitch to the 16-byte record stride
            if tiles is not None:
                dev.snp_kernel_tiles(store, chunk=chunk, tiles=tiles, accumulate=False, low_term=low_term)
                dist.all_reduce(tiles)
                dev.kernel_from_tiles(tiles, n, K=K)
            else:
                dev.snp_kernel(store, K=K, accumulate=False, chunk=chunk, mirror=False, low_term=low_term)
                dist.all_reduce(K)
                _lib.check(lib.pstb_mirror_lower(K.data_ptr(), n, n, stream))
            if rank == 0:
                t_K.copy_(K, non_blocking=True)                                     # D2H of the finished kernel
            torch.cuda.synchronize()
        api = "per rank: pinned packed shard -> HBM, pstb_snp_kernel, NCCL all-reduce, rank 0 copies float32 K to pinned host memory"
    step()
    barrier()
    steps = max(1, min(args.steps, args.kernel_steps))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    dt = max_over_ranks((time.perf_counter() - t0) / steps)
    ok = None
    if rank == 0:
        hk = np.ctypeslib.as_array(ctypes.cast(h_K, ctypes.POINTER(ctypes.c_float)), shape=(n, n))
        ok = bool(abs(float(np.mean(np.diagonal(hk).astype(np.float64))) / m - 1.0) < 1e-3 and np.array_equal(hk[:64, -64:], hk[-64:, :64].T))
        del hk
        lib.pstb_host_free(h_K)
    del h_packed
    lib.pstb_host_free(h_pk)
    lib.pstb_host_release()
    return {"value": 2.0 * n * n * m / dt / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": m * rec, "d2h_bytes_per_step": n * n * 4 + (16 * m if world == 1 else 0),
            "ms_per_step": dt * 1e3, "steps": steps, "api": api, "result_check": ok}


def run_cfg5(args, torch, dist, dev, _lib, rank, world, barrier, max_over_ranks):
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
    # parity on sampled tiles against the CPU oracle (statistics from the GPU run, themselves checked on a SNP sample)
    from oracle import bed_oracle
    rng = np.random.default_rng(rank)
    worst = 0.0
    packed_all = None
    for t in rng.choice(len(coords), size=min(args.cfg5_tiles, len(coords)), replace=False):
        I, J = int(coords[t][0]), int(coords[t][1])
        rows = np.arange(I * 256, min(n, I * 256 + 256))
        cols = np.arange(J * 256, min(n, J * 256 + 256))
        if packed_all is None:
            st = stats.cpu().numpy()
            head = store.tensor[:64, :rec].cpu().numpy()
            _sub, rst = bed_oracle.standardize(bed_oracle.decode(head, n))
            assert np.allclose(st[:64], rst, rtol=1e-12, equal_nan=True)
            packed_all = True

        def tile_rows(block):                                   # 256 aligned individuals = 64 contiguous bytes of every record
            lo_b, cnt = block * 64, min(n, block * 256 + 256) - block * 256
            sub = store.tensor[:, lo_b:lo_b + (cnt + 3) // 4].contiguous().cpu().numpy()
            x, _ = bed_oracle.standardize(bed_oracle.decode(sub, cnt), use_stats=True, stats=st)
            return x
        xr, xc = tile_rows(I), tile_rows(J)
        ref = xr @ xc.T
        got = tiles[t, : len(rows), : len(cols)].double().cpu().numpy()
        worst = max(worst, float(np.linalg.norm(got - ref) / max(1e-300, np.linalg.norm(ref))))
    worst = max_over_ranks(worst)
    return {"metric": "SnpKernel TFLOP/s (2*N^2*M)", "value": 2.0 * n * n * m / (ms * 1e-3) / 1e12, "unit": "TFLOP/s", "n_gpus": world, "ms_per_step": ms,
            "config": {"workload": "cfg5: synthetic .bed {0} iids x {1} SNPs, streamed decode+standardize+K, K-tile sharded over {2} GPUs (no collective)".format(n, m, world),
                       "tiles_per_rank": int(len(coords)), "tile_bytes_per_rank": need, "chunk_snps": chunk},
            "parity": {"sampled_tiles_per_rank": int(min(args.cfg5_tiles, len(coords))), "worst_rel_frobenius_vs_oracle": worst}}


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
    ap.add_argument("--cfg5-tiles", type=int, default=8, help="tiles per rank compared with the CPU oracle")
    ap.add_argument("--api-e2e", action="store_true", help="also time Bed(file).read(dtype=float32, standardizer=Unit()) -> NumPy through the Python layer")
    ap.add_argument("--only-kernel", action="store_true", help="experiments: run only the cfg3 SnpKernel leg and print its object")
    ap.add_argument("--no-e2e", dest="e2e", action="store_false", help="profiling runs only: skip the host-buffer leg")
    ap.add_argument("--kernel-n", type=int, default=CFG3["n_iid"])
    ap.add_argument("--kernel-m", type=int, default=CFG3["n_sid"])
    ap.add_argument("--kernel-steps", type=int, default=2)
    ap.add_argument("--kernel-chunk", type=int, default=None)
    ap.add_argument("--square-allreduce", action="store_true", help="A/B: all-reduce the square K instead of the compact lower triangle")
    ap.add_argument("--no-kernel-cpu", dest="kernel_cpu", action="store_false", help="skip the CPU SnpKernel sample (NumPy BLAS) of the kernel leg")
    ap.add_argument("--ncu-traffic", type=float, default=None, help="dram bytes per launch from profiles/ (ncu --set full), for the roofline object")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
