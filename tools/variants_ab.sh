#!/bin/bash
# Times config 4 (float32 shading unless DRT_SHADE_F32 is set otherwise) with the default library and with every variant library in
# dartray_b200/variants/ (tools: build.build(defines=..., only=...)).  Run on the GPU box.
cd "$(dirname "$0")/.."
export DRT_SHADE_F32=${DRT_SHADE_F32:-1}
one() { python tools/render_bench.py path 1920 1080 256 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  %.4f s  %.1f Msamples/s  mean %s' % (d['seconds'], d['samples_per_s']/1e6, d['mean_rgb']))"; }
for rep in 1 2; do
  echo "=== default"; one
  for lib in dartray_b200/variants/lib_*.so; do
    echo "=== $(basename $lib .so)"; DRT_LIB_PATH=$PWD/$lib one
  done
done
