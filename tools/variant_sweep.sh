#!/bin/bash
# Builds libdartray_gpu.so with different tuning macros and times the config-2 launches (run on the GPU box).
set -e
cd "$(dirname "$0")/.."
for v in "$@"; do
  mb=${v%%:*}; lb=${v##*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-pthread,-ffp-contract=off \
     -shared -cudart static -DDRT_MIN_BLOCKS=$mb -DDRT_LEAF_BATCH=$lb -o dartray_b200/libdartray_gpu.so dartray_b200/csrc/*.cu dartray_b200/csrc/*.cpp
  echo "=== MIN_BLOCKS=$mb LEAF_BATCH=$lb"
  python tools/quick_trace_bench.py 512 4194304 2>&1 | grep -E "closest:|any:"
done
