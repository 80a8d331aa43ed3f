#!/bin/bash
# Builds pysnptools_b200/libpst_b200.so for sm_100a (nvcc cross-compiles without a GPU).
set -e
here="$(cd "$(dirname "$0")" && pwd)"
out="$here/../libpst_b200.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
mkdir -p "$here/_obj"
pids=()
for f in runtime decode standardize host_api syrk syrk_f64; do
  [ -f "$here/$f.cu" ] || continue
  "$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC \
      ${PSTB_PTXAS_V:+-Xptxas -v} -c "$here/$f.cu" -o "$here/_obj/$f.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
objs=$(ls "$here"/_obj/*.o)
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$out" $objs
echo "built $out"
