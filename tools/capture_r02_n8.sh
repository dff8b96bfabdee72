# Round-2 8-GPU measurements (run with gpurun --gpus 8): the bench at N = 8 (one process per GPU, NCCL film sum), config 5 with the
# crop parity, the sharded-film check, and the same renders through ONE process and a multi-device context (drt_create_multi).
tag=${1:-r02z}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$T --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_n8.json 2> gpurun_out/${tag}_bench_n8.err
tail -c 300 gpurun_out/${tag}_bench_n8.err; head -c 300 gpurun_out/${tag}_bench_n8.json; echo
$T --master-port 29533 tools/config5.py --crop-parity > gpurun_out/${tag}_config5_n8.json 2> gpurun_out/${tag}_config5_n8.err
tail -c 300 gpurun_out/${tag}_config5_n8.err; head -c 1200 gpurun_out/${tag}_config5_n8.json; echo
$T --master-port 29541 tools/multi_gpu_check.py > gpurun_out/${tag}_multi_gpu_check.log 2>&1; tail -3 gpurun_out/${tag}_multi_gpu_check.log
python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -2 | tee gpurun_out/${tag}_multi_pytest_n8.log
python tools/multi_ctx_bench.py --config5 --spp5 1024 2>&1 | tee gpurun_out/${tag}_multi_ctx_bench_n8.log
