# A/B of the shape-class sort in front of the matte-only path kernel (run on the GPU box)
cd "$(dirname "$0")/.."
for rep in 1 2; do
echo "=== shape sort (DRT_SHAPE_SORT=1)"; DRT_SHAPE_SORT=1 python tools/render_bench.py path 1920 1080 256 2>&1 | tail -1 | cut -c1-230
echo "=== queue order (default)"; python tools/render_bench.py path 1920 1080 256 2>&1 | tail -1 | cut -c1-230
done
echo "=== soup 512 spheres, shape sort"; DRT_SHAPE_SORT=1 python tools/render_bench.py soup 1920 1080 16 512 2>&1 | tail -1 | cut -c1-230
echo "=== soup 512 spheres, queue order"; python tools/render_bench.py soup 1920 1080 16 512 2>&1 | tail -1 | cut -c1-230
