# Round-2 closing capture (run under gpurun): everything tools/capture_r02.sh takes, plus the float32-shading path render (launch
# list, --set full of its path-vertex kernel and of the small-scene traversal kernel).  Outputs in gpurun_out/<tag>_*; tag = $1
tag=${1:-r02final}
bash tools/capture_r02.sh $tag
DRT_SHADE_F32=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${tag}_path_f32_launches.csv python tools/render_bench.py path 960 540 64 > gpurun_out/${tag}_ncu_path_f32.log 2>&1
DRT_SHADE_F32=1 ncu --set full --clock-control none --import-source on -k regex:shadePathKernel -s 7 -c 1 -o gpurun_out/${tag}_shade_f32 -f python tools/render_bench.py path 960 540 64 > /dev/null 2>&1
DRT_SHADE_F32=1 ncu --set full --clock-control none --import-source on -k regex:traceSmallKernel -s 3 -c 1 -o gpurun_out/${tag}_small -f python tools/render_bench.py path 960 540 64 > /dev/null 2>&1
for k in shade_f32 small; do
  python tools/ncu_metrics.py gpurun_out/${tag}_$k.ncu-rep > gpurun_out/${tag}_${k}_metrics.txt 2>&1
  python tools/ncu_hot_lines.py gpurun_out/${tag}_$k.ncu-rep 0 60 > gpurun_out/${tag}_${k}_lines.txt 2>&1
done
echo "--- path f64 launches"; python tools/launch_sum.py gpurun_out/${tag}_path_launches.csv | head -12
echo "--- path f32 launches"; python tools/launch_sum.py gpurun_out/${tag}_path_f32_launches.csv | head -12
