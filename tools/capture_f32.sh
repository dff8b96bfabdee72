# ncu evidence for the float32 path build (run under gpurun): launch list of a config-4-style render in both precisions, --set full of
# the float32 path-vertex kernel and of the low-discrepancy sampler kernel.  tag = $1
tag=${1:-r02z}
cd "$(dirname "$0")/.."
DRT_SHADE_F32=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${tag}_path_f32_launches.csv python tools/render_bench.py path 960 540 64 > gpurun_out/${tag}_ncu_path_f32.log 2>&1
DRT_SHADE_F32=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${tag}_path_f64_launches.csv python tools/render_bench.py path 960 540 64 > gpurun_out/${tag}_ncu_path_f64.log 2>&1
DRT_SHADE_F32=1 ncu --set full --clock-control none --import-source on -k regex:shadePathKernel -s 7 -c 1 -o gpurun_out/${tag}_shade_f32 -f python tools/render_bench.py path 960 540 64 > gpurun_out/${tag}_ncu_shade_f32.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:samplerLD -s 2 -c 1 -o gpurun_out/${tag}_sampler -f python tools/render_bench.py path 960 540 64 > gpurun_out/${tag}_ncu_sampler.log 2>&1
python tools/ncu_metrics.py gpurun_out/${tag}_shade_f32.ncu-rep > gpurun_out/${tag}_shade_f32_metrics.txt 2>&1
python tools/ncu_metrics.py gpurun_out/${tag}_sampler.ncu-rep > gpurun_out/${tag}_sampler_metrics.txt 2>&1
python tools/ncu_hot_lines.py gpurun_out/${tag}_shade_f32.ncu-rep > gpurun_out/${tag}_shade_f32_lines.txt 2>&1
python tools/ncu_hot_lines.py gpurun_out/${tag}_sampler.ncu-rep > gpurun_out/${tag}_sampler_lines.txt 2>&1
python tools/launch_sum.py gpurun_out/${tag}_path_f32_launches.csv 2>&1 | head -30
python tools/launch_sum.py gpurun_out/${tag}_path_f64_launches.csv 2>&1 | head -30
