# 8-GPU measurements of a round (run with gpurun --gpus 8): bench at N = 8, config 5, sharded-film check
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$T --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/o_bench_n8.json 2> gpurun_out/o_bench_n8.err
tail -c 300 gpurun_out/o_bench_n8.err; head -c 400 gpurun_out/o_bench_n8.json; echo
$T --master-port 29533 tools/config5.py --crop-parity > gpurun_out/o_config5_n8.json 2> gpurun_out/o_config5_n8.err
tail -c 300 gpurun_out/o_config5_n8.err; head -c 900 gpurun_out/o_config5_n8.json; echo
$T --master-port 29541 tools/multi_gpu_check.py > gpurun_out/o_multi_gpu_check.log 2>&1; tail -3 gpurun_out/o_multi_gpu_check.log
