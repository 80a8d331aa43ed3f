#!/bin/bash
# Multi-GPU checks (gpurun --gpus N -- bash scripts/run_multi_gpu.sh N [kernel_m]): the SNP-sharded SnpKernel leg with the overlapped and
# with the serial (round-1) all-reduce, then the whole default bench line.
N=${1:-2}
M=${2:-200000}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
mkdir -p gpurun_out
$RUN bench.py --gpus $N --only-kernel --kernel-m $M --no-kernel-cpu --kernel-steps 3 --steps 3 > gpurun_out/k${N}_overlap.json 2> gpurun_out/k${N}_overlap.err; echo "rc=$?" >> gpurun_out/k${N}_overlap.err
$RUN bench.py --gpus $N --only-kernel --kernel-m $M --no-kernel-cpu --kernel-steps 3 --steps 3 --serial-allreduce --no-e2e --kernel-parity-blocks 0 > gpurun_out/k${N}_serial.json 2> gpurun_out/k${N}_serial.err; echo "rc=$?" >> gpurun_out/k${N}_serial.err
if [ "${3:-full}" = "full" ]; then
$RUN bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?" >> gpurun_out/bench_${N}gpu.err
fi
