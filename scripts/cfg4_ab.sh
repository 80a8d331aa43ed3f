for r in 1 2 3; do
python scripts/bench_read.py cfg4real 2>&1 | sed 's/^/pool /'
done
PSTB_HOST_TRACE=1 python scripts/bench_pageable.py
