# One round's 1-GPU capture (run under gpurun): tests, smoke, bench (both arms), ncu launch lists.  Outputs in gpurun_out/<tag>_*.
tag=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 300 gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-render > gpurun_out/${tag}_ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${tag}_path_launches.csv python tools/render_bench.py path 960 540 64 > gpurun_out/${tag}_ncu_path.log 2>&1
head -c 300 gpurun_out/${tag}_bench.json; echo
