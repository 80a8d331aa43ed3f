# round-end validation on one B200: GPU tests, the default bench line, the reference arm; with "ncu" as first argument also the launch
# list of the bench command and one full ncu capture of k_read_f (profiles/r2_bench_launches.csv, r2_read_f_full.txt)
python -m pytest tests -q -m gpu -x > gpurun_out/t_final.log 2>&1; echo "rc=$?" >> gpurun_out/t_final.log
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc=$?" >> gpurun_out/t_final.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
if [ "$1" = "ncu" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_bench_launches.csv \
      python bench.py --steps 3 --warmup 3 --no-e2e --no-api-e2e --no-extra-legs --kernel-steps 1 --no-kernel-cpu --kernel-parity-blocks 0 > gpurun_out/ncu_bench.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_read_f -c 1 -f -o gpurun_out/r2_read_f python scripts/bench_read.py cfg2 > gpurun_out/ncu_read_f.log 2>&1
fi
