#!/bin/bash
# A/B of the wavefront batch size and of the sampler kernel's shared-memory budget on config 4 with float32 shading.  Run on the GPU box.
cd "$(dirname "$0")/.."
export DRT_SHADE_F32=1
one() { python tools/render_bench.py path 1920 1080 256 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  %.4f s  %.1f Msamples/s' % (d['seconds'], d['samples_per_s']/1e6))"; }
for rep in 1 2; do
  echo "=== default (16 Mi slots, 48 KB)"; one
  for s in 4194304 8388608 33554432; do echo "=== SLOTS=$s"; SLOTS=$s one; done
  for kb in 32 72 100 140; do echo "=== DRT_SAMPLER_SMEM_KB=$kb"; DRT_SAMPLER_SMEM_KB=$kb one; done
done
