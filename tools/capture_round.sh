set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; tail -c 600 gpurun_out/n_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/n_bench_ref.json 2>> gpurun_out/n_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/n_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-render > gpurun_out/n_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:traceFastKernel -s 3 -c 3 -o gpurun_out/n_full -f python tools/prof_one.py > gpurun_out/n_ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/n_path_launches.csv python tools/render_bench.py path 960 540 64 > gpurun_out/n_ncu_path.log 2>&1
head -c 1500 gpurun_out/n_bench.json
