# 8-GPU round-end run (gpurun --gpus 8 -- bash scripts/final_8gpu.sh 8): the default bench line under torchrun, the host-bandwidth probe, and the cfg2 legs without the NUMA binding (A/B)
N=${1:-8}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$RUN bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?" >> gpurun_out/bench_${N}gpu.err
timeout 300 python scripts/probe_host_bw.py > gpurun_out/host_bw_${N}gpu.txt 2>&1
$RUN bench.py --gpus $N --steps 5 --warmup 3 --no-kernel --no-extra-legs --no-api-e2e --no-numa-bind > gpurun_out/bench_${N}gpu_nobind.json 2> gpurun_out/bench_${N}gpu_nobind.err
