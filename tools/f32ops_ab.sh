cd /root/repo
for rep in 1 2; do
echo "=== default (f32 single ops)"; python tools/render_bench.py path 1920 1080 256 2>&1 | tail -1 | cut -c1-300
echo "=== f64ops"; DRT_LIB_PATH=$PWD/dartray_b200/variants/lib_f64ops.so python tools/render_bench.py path 1920 1080 256 2>&1 | tail -1 | cut -c1-300
done
echo "=== materials default"; python tools/render_bench.py materials 1920 1080 64 2>&1 | tail -1 | cut -c1-300
echo "=== materials f64ops"; DRT_LIB_PATH=$PWD/dartray_b200/variants/lib_f64ops.so python tools/render_bench.py materials 1920 1080 64 2>&1 | tail -1 | cut -c1-300
