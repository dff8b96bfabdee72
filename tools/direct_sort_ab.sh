# A/B of material order for directSampleKernel now that the sort costs 0.1 ms (DRT_DIRECT_SORT), cornell_materials through the
# directlighting integrator; parity tests of the direct-lighting path under the knob
mkdir -p gpurun_out
DRT_DIRECT_SORT=1 python -m pytest tests/test_render_gpu.py -m gpu -x -q -k "direct or bxdf_lists" 2>&1 | tail -2
for i in 1 2; do
  INTEG=direct python tools/render_bench.py materials 1920 1080 16 | tail -1 | head -c 160; echo " [queue order]"
  DRT_DIRECT_SORT=1 INTEG=direct python tools/render_bench.py materials 1920 1080 16 | tail -1 | head -c 160; echo " [material order]"
done
python tools/render_bench.py materials 1920 1080 64 | tail -1 | head -c 160; echo " [path, unchanged]"
