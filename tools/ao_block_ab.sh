#!/bin/bash
# A/B of the AO queue's transposed block size (variant libraries built with -DDRT_AO_BLOCK_LOG2=n).  Run on the GPU box.
cd "$(dirname "$0")/.."
one() { python tools/render_bench.py ao 1920 1080 1 64 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  soup_1m %.5f s' % d['seconds'])"; INTEG=ao python tools/render_bench.py path 1920 1080 4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  cornell %.5f s' % d['seconds'])"; }
for rep in 1 2; do
  echo "=== default (32 hits)"; one
  for lib in dartray_b200/variants/lib_aoblk*.so; do echo "=== $(basename $lib .so)"; DRT_LIB_PATH=$PWD/$lib one; done
done
