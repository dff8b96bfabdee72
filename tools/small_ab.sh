#!/bin/bash
# A/B of the leaf-list kernel for small scenes (traceSmallKernel; DRT_NO_SMALL=1 keeps the persistent float32-node kernel) on config 4,
# in both shading precisions.  Run on the GPU box.
cd "$(dirname "$0")/.."
one() { python tools/render_bench.py path 1920 1080 256 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  %.4f s  %.1f Msamples/s  mean %s' % (d['seconds'], d['samples_per_s']/1e6, d['mean_rgb']))"; }
for rep in 1 2; do
  echo "=== f64, tree kernel"; DRT_NO_SMALL=1 DRT_SHADE_F32=0 one
  echo "=== f64, leaf-list kernel"; DRT_SHADE_F32=0 one
  echo "=== f32, tree kernel"; DRT_NO_SMALL=1 DRT_SHADE_F32=1 one
  echo "=== f32, leaf-list kernel"; DRT_SHADE_F32=1 one
done
