#!/bin/bash
# Builds libdartray_gpu.so with different shading-kernel tuning macros and times config 4 at 16 spp (run on the GPU box).
cd "$(dirname "$0")/.."
for v in "$@"; do
  mb=${v%%:*}; inl=${v##*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-pthread,-ffp-contract=off \
     -shared -cudart static -DDRT_SHADE_MIN_BLOCKS=$mb -DDRT_SHAPE_INLINE=$inl -Xptxas -v -o dartray_b200/libdartray_gpu.so dartray_b200/csrc/*.cu dartray_b200/csrc/*.cpp 2>&1 | grep -A2 "shadePathKernel" | grep -E "registers|spill" 
  echo "=== SHADE_MIN_BLOCKS=$mb SHAPE_INLINE=$inl"
  python tools/render_bench.py path 1920 1080 16 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['seconds'], d['samples_per_s']/1e6, 'Msamples/s')"
done
