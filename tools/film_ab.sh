# A/B of the film kernel's warp sum (DRT_FILM_PLAIN_ATOMICS = the old four atomics per sample) on config 4, then the full GPU suite
mkdir -p gpurun_out
for i in 1 2; do
  python tools/render_bench.py path 1920 1080 256 | tail -1 | head -c 330; echo " [warp sum]"
  DRT_FILM_PLAIN_ATOMICS=1 python tools/render_bench.py path 1920 1080 256 | tail -1 | head -c 330; echo " [per sample]"
done
python -m pytest tests -m gpu -x -q > gpurun_out/y_pytest.log 2>&1; tail -2 gpurun_out/y_pytest.log
