# A/B of material order for whittedSampleKernel (DRT_WHITTED_SORT) on cornell_materials, after the full GPU suite on the new defaults
# (directSampleKernel in material order)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/v_pytest.log 2>&1; tail -2 gpurun_out/v_pytest.log
DRT_WHITTED_SORT=1 python -m pytest tests/test_render_gpu.py -m gpu -x -q -k "whitted" 2>&1 | tail -2
for i in 1 2; do
  INTEG=whitted python tools/render_bench.py materials 1920 1080 16 | tail -1 | head -c 160; echo " [queue order]"
  DRT_WHITTED_SORT=1 INTEG=whitted python tools/render_bench.py materials 1920 1080 16 | tail -1 | head -c 160; echo " [material order]"
done
INTEG=direct python tools/render_bench.py materials 1920 1080 16 | tail -1 | head -c 160; echo " [direct, default]"
