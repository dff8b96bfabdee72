#!/bin/bash
# A/B of the AO ray queue layout (cell-major blocks of 32 hits against the reference's hit-major order) on config 3.  Run on the GPU box.
cd "$(dirname "$0")/.."
one() { python tools/render_bench.py ao 1920 1080 1 64 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  %.5f s  %.1f Mrays/s  mean %s' % (d['seconds'], d['mrays_per_s'], d['mean_rgb']))"; }
for rep in 1 2 3; do
  echo "=== hit-major (DRT_AO_PLAIN_ORDER=1)"; DRT_AO_PLAIN_ORDER=1 one
  echo "=== cell-major blocks (default)"; one
done
