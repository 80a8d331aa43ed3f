#!/bin/bash
# Multi-GPU checks (gpurun --gpus N -- bash scripts/run_multi_gpu.sh N [full|kernel] [kernel_m]): the whole default bench line under torchrun, or
# only the SNP-sharded SnpKernel leg with the default (sliced, serial) and the band-overlapped all-reduce.
N=${1:-2}
MODE=${2:-full}
M=${3:-500000}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
mkdir -p gpurun_out
if [ "$MODE" = "full" ]; then
  $RUN bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?" >> gpurun_out/bench_${N}gpu.err
else
  KARGS="--gpus $N --only-kernel --kernel-m $M --no-kernel-cpu --kernel-steps 3 --steps 3 --no-e2e --kernel-parity-blocks 0"
  $RUN bench.py $KARGS --serial-allreduce > gpurun_out/k${N}_serial.json 2> gpurun_out/k${N}_default.err; echo "rc=$?" >> gpurun_out/k${N}_default.err
  PSTB_OVERLAP_TRACE=1 $RUN bench.py $KARGS --allreduce-sms 0 > gpurun_out/k${N}_overlap_s0.json 2> gpurun_out/k${N}_overlap_s0.err
  $RUN bench.py $KARGS --allreduce-sms 0 --allreduce-bands 4 --kernel-parity-blocks 11 > gpurun_out/k${N}_overlap_s0_b4.json 2> gpurun_out/k${N}_overlap_s0_b4.err
  if [ "$N" -le 2 ]; then
    $RUN bench.py $KARGS --serial-allreduce --allreduce-slices 1 > gpurun_out/k${N}_slices1.json 2> gpurun_out/k${N}_slices1.err
    $RUN bench.py $KARGS > gpurun_out/k${N}_overlap.json 2> gpurun_out/k${N}_overlap.err
    PSTB_SYRK_DYN=0 $RUN bench.py $KARGS > gpurun_out/k${N}_overlap_static.json 2> gpurun_out/k${N}_overlap_static.err
  fi
fi
