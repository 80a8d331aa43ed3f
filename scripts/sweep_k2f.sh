#!/bin/bash
# tuning sweep of the staged K2f kernel (threads per CTA x buffers per CTA)
for t in 128 256 512; do for nb in 1 2; do
  echo "threads=$t nbuf=$nb"; PSTB_STD_THREADS=$t PSTB_STD_NBUF=$nb python scripts/bench_read.py k2f 2>&1 | grep "K2f.* F "
done; done
