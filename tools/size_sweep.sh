#!/bin/bash
# Quantised-node kernel against the float32-node kernel over scene sizes (run on the GPU box).
cd "$(dirname "$0")/.."
for ns in 1 2 4 8 64; do
  echo "=== soup($ns) FAST_Q"; DRT_Q_MIN_PRIMS=0 python tools/quick_trace_bench.py $ns 4194304 2>&1 | grep -E "closest:|any:"
  echo "=== soup($ns) FAST_V1"; DRT_TRACE_V1=1 python tools/quick_trace_bench.py $ns 4194304 2>&1 | grep -E "closest:|any:"
done
