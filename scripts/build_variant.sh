#!/bin/bash
# build an experimental variant of the library: scripts/build_variant.sh <name> <extra nvcc flags...>
set -e
root="$(cd "$(dirname "$0")/.." && pwd)"
name="$1"; shift
mkdir -p "$root/gpurun_variants"
objs=""
for f in runtime decode standardize host_api syrk; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c "$root/pysnptools_b200/csrc/$f.cu" -o "/tmp/var_${name}_$f.o" &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$root/gpurun_variants/libpst_$name.so" /tmp/var_${name}_*.o
echo "built gpurun_variants/libpst_$name.so"
