python -m pytest tests/test_gpu_read.py tests/test_gpu_api.py tests/test_gpu_kernel.py -q -m gpu -x > gpurun_out/t22.log 2>&1; echo "rc=$?" >> gpurun_out/t22.log
python scripts/prof_api_read.py > gpurun_out/prof_api3.txt 2>&1
python scripts/bench_read.py feed_i8 > gpurun_out/feed_i8.txt 2>&1
