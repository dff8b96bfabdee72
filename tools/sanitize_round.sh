# compute-sanitizer memcheck over the GPU tests of this session's features (quadrics, mesh attributes, infinite light, halton)
# and a 2-rank bench to confirm the multi-GPU path (run with gpurun --gpus 2)
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_trace_gpu.py tests/test_render_gpu.py -m gpu -x -q \
  -k "quadrics or cylinders or smooth or infinite or halton or cone_light" > gpurun_out/n_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/n_memcheck.log
tail -5 gpurun_out/n_memcheck.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/n_bench_n2.json 2> gpurun_out/n_bench_n2.err
tail -c 300 gpurun_out/n_bench_n2.err; head -c 700 gpurun_out/n_bench_n2.json
