# compute-sanitizer over the kernels of the round's last session: the leaf-list traversal kernel (edge-case test), the float32 path
# kernels, the rewritten low-discrepancy sampler kernel (shared-memory tables).  Run under gpurun; logs in gpurun_out/r02final_sanitizer_*.log
mkdir -p gpurun_out
compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_trace_gpu.py -m gpu -q -x -k "small_scene or (default and (known_answer or big_leaf or axis_aligned))" > gpurun_out/r02final_sanitizer_memcheck_small.log 2>&1; echo "memcheck small rc=$?"
compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_precision_gpu.py tests/test_config1.py -m gpu -q -x -k "small_film or bxdf_list or validated or keyed_oracle" > gpurun_out/r02final_sanitizer_memcheck_f32.log 2>&1; echo "memcheck f32 rc=$?"
compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_precision_gpu.py tests/test_trace_gpu.py -m gpu -q -x -k "bxdf_list or (default and known_answer)" > gpurun_out/r02final_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_precision_gpu.py tests/test_trace_gpu.py -m gpu -q -x -k "bxdf_list or (default and axis_aligned)" > gpurun_out/r02final_sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?"
for f in gpurun_out/r02final_sanitizer_*.log; do tail -n 4 $f; done
# the float32 traversal units (lowered trace_fast.cu / trace_fast2.cu) through the float32 path render
compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_precision_gpu.py -m gpu -q -x -k "16k or bxdf_list or (mesh_attributes and matte)" > gpurun_out/r02final_sanitizer_memcheck_f32trace.log 2>&1; echo "memcheck f32 trace rc=$?"
compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_precision_gpu.py -m gpu -q -x -k "16k" > gpurun_out/r02final_sanitizer_racecheck_f32trace.log 2>&1; echo "racecheck f32 trace rc=$?"
# the AO generator with its shared-memory inverse table
compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_render_gpu.py -m gpu -q -x -k "ambient" > gpurun_out/r02final_sanitizer_memcheck_ao.log 2>&1; echo "memcheck ao rc=$?"
compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_render_gpu.py -m gpu -q -x -k "ray_queue_layout and 64" > gpurun_out/r02final_sanitizer_racecheck_ao.log 2>&1; echo "racecheck ao rc=$?"
