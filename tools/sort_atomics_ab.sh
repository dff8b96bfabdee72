# A/B of the material counting sort's atomics (warp-aggregated vs one per queue entry) on cornell_materials, the BxDF-list parity
# tests on the new kernels, and a launch list for the share of the sort passes
mkdir -p gpurun_out
python -m pytest tests/test_render_gpu.py -m gpu -x -q -k "bxdf_lists or translucent or whitted" 2>&1 | tail -2
for i in 1 2; do
  python tools/render_bench.py materials 1920 1080 64 | tail -1 | head -c 400; echo " [aggregated]"
  DRT_SORT_PLAIN_ATOMICS=1 python tools/render_bench.py materials 1920 1080 64 | tail -1 | head -c 400; echo " [plain]"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/t_materials_launches.csv python tools/render_bench.py materials 960 540 16 > gpurun_out/t_ncu_materials.log 2>&1
DRT_SORT_PLAIN_ATOMICS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/t_materials_launches_plain.csv python tools/render_bench.py materials 960 540 16 > gpurun_out/t_ncu_materials_plain.log 2>&1
