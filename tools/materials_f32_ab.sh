cd /root/repo
one() { python tools/render_bench.py materials 1920 1080 64 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  %.4f s  %.1f Msamples/s  mean %s' % (d['seconds'], d['samples_per_s']/1e6, d['mean_rgb']))"; }
for rep in 1 2; do echo "=== materials f64"; DRT_SHADE_F32=0 one; echo "=== materials f32"; DRT_SHADE_F32=1 one; done
