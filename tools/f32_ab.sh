#!/bin/bash
# A/B of the float32 shading build (DRT_SHADE_F32=1) against the binary64 kernels on config 4, and of the float32 unit's occupancy
# variants (dartray_b200/variants/lib_f32mb*.so, built by hand with -DDRT_SHADE_MIN_BLOCKS=n for render_kernels_f32.cu).  Run on the GPU box.
cd "$(dirname "$0")/.."
one() { python tools/render_bench.py path 1920 1080 256 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  %.4f s  %.1f Msamples/s  mean %s' % (d['seconds'], d['samples_per_s']/1e6, d['mean_rgb']))"; }
for rep in 1 2; do
  echo "=== binary64 (default)"; DRT_SHADE_F32=0 one
  echo "=== float32, min blocks 4 (128 regs)"; DRT_SHADE_F32=1 one
  for v in f32mb5 f32mb6 f32mb8; do
    if [ -f dartray_b200/variants/lib_$v.so ]; then echo "=== float32 $v"; DRT_LIB_PATH=$PWD/dartray_b200/variants/lib_$v.so DRT_SHADE_F32=1 one; fi
  done
done
