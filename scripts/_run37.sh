python -m pytest tests/test_gpu_read.py tests/test_gpu_api.py -q -m gpu -x > gpurun_out/t23.log 2>&1; echo "rc=$?" >> gpurun_out/t23.log
python scripts/prof_api_read.py > gpurun_out/prof_api4.txt 2>&1
python scripts/bench_pageable.py 500000 > gpurun_out/pageable4.txt 2>&1
python scripts/bench_read.py variants > gpurun_out/variants4.txt 2>&1
