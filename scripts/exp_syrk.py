"""K3 tuning experiments on a cfg3-shaped problem (N = 50 000, a few SNP chunks): time pstb_snp_kernel under the environment knobs of
csrc/syrk.cu (PSTB_SYRK_RED, PSTB_SYRK_DBG, PSTB_RUN_KB_FAST, PSTB_SYRK_GROUP, PSTB_SYRK_CLUSTERS) and different chunk sizes.
Usage: python scripts/exp_syrk.py [N] [M] [missing_rate]   -> one line per variant: ms, TFLOP/s (2 N^2 M)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from pysnptools_b200 import device as dev

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 8 * 4032
missing = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
store = bench.gen_store_device(dev, torch, n, m, seed=2000, missing_rate=missing)
K = torch.zeros((n, n), dtype=torch.float32, device="cuda")
KNOBS = ("PSTB_SYRK_DYN", "PSTB_RUN_KB_OFF", "PSTB_SYRK_TMA_OUT", "PSTB_SYRK_RED", "PSTB_SYRK_DBG", "PSTB_RUN_KB_FAST", "PSTB_SYRK_GROUP", "PSTB_SYRK_CLUSTERS", "PSTB_SYRK_3TERM")


sampler = bench.ClockSampler(0)
sampler.wait_started()
import time


def run(label, chunk=4032, low_term="fp8", reps=3, **env):
    for k in KNOBS:
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ[k] = str(v)
    dev.snp_kernel(store, K=K, accumulate=False, chunk=chunk, mirror=False, low_term=low_term)
    torch.cuda.synchronize()
    best = 1e30
    t0 = time.perf_counter()
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dev.snp_kernel(store, K=K, accumulate=False, chunk=chunk, mirror=False, low_term=low_term)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    ck = sampler.window(t0, time.perf_counter())
    print("%-58s chunk %5d  %8.2f ms  %7.1f TFLOP/s   sm %s MHz %s W" % (label, chunk, best, 2.0 * n * n * m / best / 1e9, ck.get("sm_mhz"), ck.get("power_w_max")), flush=True)
    for k in KNOBS:
        os.environ.pop(k, None)


print("N = %d, M = %d, missing = %.2f" % (n, m, missing), flush=True)
import ctypes
from pysnptools_b200 import _lib
probe = _lib.lib.pstb_debug_max_active_clusters
probe.restype = ctypes.c_int
for cs in (1, 2, 4, 8):
    print("max active clusters of %d CTAs (320 threads, 225 KB smem): %d" % (cs, probe(cs, 320, 225 * 1024)), flush=True)
run("default (dynamic tile feed, TMA store / reduce-add epilogue, fp8 low term)")
run("static round-robin tile lists", PSTB_SYRK_DYN=0)
run("default again")
run("static again", PSTB_SYRK_DYN=0)
for rk in (6, 12, 24, 63):
    run("off-diagonal tiles: %d k-blocks per TMEM run" % rk, PSTB_RUN_KB_OFF=rk)
    run("off-diagonal tiles: %d k-blocks per TMEM run, fp16 low term" % rk, low_term="fp16", PSTB_RUN_KB_OFF=rk)
run("per-thread red.add epilogue", PSTB_SYRK_TMA_OUT=0)
run("round-1 epilogue (load + add + store)", PSTB_SYRK_TMA_OUT=0, PSTB_SYRK_RED=0)
run("no K write at all [timing only]", PSTB_SYRK_DBG=2)
run("no operand loads [timing only]", PSTB_SYRK_DBG=4)
run("no loads, no K write [timing only]", PSTB_SYRK_DBG=6)
run("fp16 low term", low_term="fp16")
run("fp16 low term, round-1 epilogue", low_term="fp16", PSTB_SYRK_TMA_OUT=0, PSTB_SYRK_RED=0)
run("3-term split", PSTB_SYRK_3TERM=1)
for chunk in (2048, 8064, 16128):
    run("default", chunk=chunk)
    run("round-1 epilogue", chunk=chunk, PSTB_SYRK_TMA_OUT=0, PSTB_SYRK_RED=0)
for rk in (3, 12, 24):
    run("run_kb_fast = %d" % rk, PSTB_RUN_KB_FAST=rk)
for g in (4, 6, 12, 16):
    run("super-block %d x %d tiles" % (g, g), PSTB_SYRK_GROUP=g)
    run("super-block %d x %d tiles" % (g, g), chunk=8064, PSTB_SYRK_GROUP=g)
for c in (72, 70, 64):
    run("%d CTA pairs" % c, PSTB_SYRK_CLUSTERS=c)
sampler.close()
