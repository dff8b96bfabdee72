# compute-sanitizer over the code this round added: the quantised-node traversal kernel (incl. its edge-case tests), the volume
# kernels, the multi-device context, the render profile.  Run under gpurun; logs in gpurun_out/r02_sanitizer_*.log
mkdir -p gpurun_out
SEL='fast_q and (random_soup or mixed or extreme or unnormalised or ties or zero) or concurrent'
compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_trace_gpu.py -m gpu -q -x -k "$SEL" > gpurun_out/r02_sanitizer_memcheck_trace.log 2>&1; echo "memcheck trace rc=$?"
compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_volumes.py tests/test_multi_gpu.py -m gpu -q -x -k "homogeneous and path or volumegrid and single and direct or aggregate or filter_footprints or shares_one" > gpurun_out/r02_sanitizer_memcheck_render.log 2>&1; echo "memcheck render rc=$?"
compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_trace_gpu.py -m gpu -q -x -k "fast_q and (mixed or extreme)" > gpurun_out/r02_sanitizer_racecheck_trace.log 2>&1; echo "racecheck trace rc=$?"
compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_trace_gpu.py -m gpu -q -x -k "fast_q and mixed" > gpurun_out/r02_sanitizer_synccheck_trace.log 2>&1; echo "synccheck trace rc=$?"
tail -3 gpurun_out/r02_sanitizer_*.log
