# compute-sanitizer over the GPU tests of the features added after capture n (adaptive / bestcandidate samplers, mapped lights,
# wrapped and substrate materials, BxDF lists with the per-bounce material counting sort): memcheck over all of them, racecheck
# over the BxDF-list path test (matHistKernel keeps its histogram in shared memory)
mkdir -p gpurun_out
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_render_gpu.py -m gpu -x -q \
  -k "adaptive or translucent or projection or best_candidate or bxdf_lists or sequence_samplers or substrate" > gpurun_out/s_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/s_memcheck.log
tail -5 gpurun_out/s_memcheck.log
timeout 150 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_render_gpu.py -m gpu -x -q \
  -k "path_integrator_with_bxdf_lists" > gpurun_out/s_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/s_racecheck.log
tail -5 gpurun_out/s_racecheck.log
