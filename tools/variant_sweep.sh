#!/bin/bash
# Builds libdartray_gpu.so with different tuning macros and times the config-2 launches (run on the GPU box).
# usage: tools/variant_sweep.sh "-DDRT_MIN_BLOCKS=5 -DDRT_LEAF_BATCH=16" "-DRAY_CHUNK=64" ...
cd "$(dirname "$0")/.."
for v in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-pthread,-ffp-contract=off \
     -shared -cudart static $v -o dartray_b200/libdartray_gpu.so dartray_b200/csrc/*.cu dartray_b200/csrc/*.cpp || exit 1
  echo "=== $v"
  python tools/quick_trace_bench.py 512 ${NRAYS:-4194304} 2>&1 | grep -E "closest:|any:"
done
