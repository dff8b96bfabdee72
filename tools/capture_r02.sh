# Round-2 1-GPU capture (run under gpurun): tests, smoke, bench (both arms), ncu launch list of the bench command, ncu --set full of
# the dominant kernel on the bench's own launch size (and the roofline traffic file read from it), path-render launch list.
# Outputs in gpurun_out/<tag>_*;  tag = $1 (default r02z)
tag=${1:-r02z}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/${tag}_smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench.err
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2>> gpurun_out/${tag}_bench.err; tail -c 300 gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-render > gpurun_out/${tag}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:traceQKernel -s 1 -c 1 -o gpurun_out/${tag}_q_closest -f python tools/prof_one.py 8388608 closest > gpurun_out/${tag}_ncu_full.log 2>&1
python tools/ncu_metrics.py gpurun_out/${tag}_q_closest.ncu-rep > gpurun_out/${tag}_ncu_full_metrics.txt 2>&1
python - <<PY
import csv, io, subprocess, json
out = subprocess.run(['ncu','-i','gpurun_out/${tag}_q_closest.ncu-rep','--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr, units, d = rows[0], rows[1], rows[2]
def val(name):
    i = hdr.index(name); v = float(d[i]); u = units[i]
    return v * {'Mbyte':1e6,'Gbyte':1e9,'Kbyte':1e3,'byte':1}.get(u,1)
rd, wr = val('dram__bytes_read.sum'), val('dram__bytes_write.sum')
json.dump({"kernel": "traceQKernel<closest, QUAD=0> on the config-2 incoherent set, 8,388,608 rays (the bench's own launch)",
           "dram_bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_written": int(wr),
           "source": "profiles/${tag}_ncu_full_metrics.txt (ncu --set full --clock-control none on tools/prof_one.py 8388608 closest, tools/capture_r02.sh)"},
          open('gpurun_out/${tag}_roofline_traffic.json','w'))
print(open('gpurun_out/${tag}_roofline_traffic.json').read())
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${tag}_path_launches.csv python tools/render_bench.py path 960 540 64 > gpurun_out/${tag}_ncu_path.log 2>&1
head -c 400 gpurun_out/${tag}_bench.json; echo
