#!/bin/bash
# Times the config-2 launches with every library in dartray_b200/variants/ (run on the GPU box, after tools/build_variants.py).
cd "$(dirname "$0")/.."
echo "=== default"; python tools/quick_trace_bench.py 512 ${NRAYS:-8388608} 2>&1 | grep -E "closest:|any:"
echo "=== default, DRT_TRACE_V1"; DRT_TRACE_V1=1 python tools/quick_trace_bench.py 512 ${NRAYS:-8388608} 2>&1 | grep -E "closest:|any:"
for lib in dartray_b200/variants/lib_*.so; do
  echo "=== $lib"
  DRT_LIB_PATH=$PWD/$lib python tools/quick_trace_bench.py 512 ${NRAYS:-8388608} 2>&1 | grep -E "closest:|any:"
done
