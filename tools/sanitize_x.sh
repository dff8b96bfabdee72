# compute-sanitizer over the warp-aggregated sort passes and the sorted direct-lighting stage (BxDF-list and direct-lighting tests)
mkdir -p gpurun_out
timeout 120 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_render_gpu.py -m gpu -x -q -k "bxdf_lists or direct" > gpurun_out/x_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/x_memcheck.log; tail -4 gpurun_out/x_memcheck.log
timeout 120 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_render_gpu.py -m gpu -x -q -k "path_integrator_with_bxdf_lists or glossy_bxdf" > gpurun_out/x_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/x_racecheck.log; tail -4 gpurun_out/x_racecheck.log
timeout 60 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_render_gpu.py -m gpu -x -q -k "path_integrator_with_bxdf_lists" > gpurun_out/x_synccheck.log 2>&1
echo "synccheck exit $?" >> gpurun_out/x_synccheck.log; tail -4 gpurun_out/x_synccheck.log
