#!/bin/bash
# A/B of traceQKernel's any-hit tuning on the AO rays of config 3 (coherent since the queue is cell-major) and on config 2's any-hit set.
cd "$(dirname "$0")/.."
one() { python tools/render_bench.py ao 1920 1080 1 64 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  config 3 %.5f s' % d['seconds'])"; python tools/quick_trace_bench.py 2>&1 | grep -E "^(incoherent|coherent) (closest|any):" | tr "\n" ";"; echo; }
echo "=== default"; one
for lib in dartray_b200/variants/lib_*.so; do echo "=== $(basename $lib .so)"; DRT_LIB_PATH=$PWD/$lib one; done
