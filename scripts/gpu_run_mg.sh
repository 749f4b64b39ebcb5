mkdir -p gpurun_out
N=${1:-2}
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/r2_dist_check_n$N.log 2>&1; grep -E "dist_check|NCCL INFO (Using|comm .* Init COMPLETE|Connected all|NVLS)" gpurun_out/r2_dist_check_n$N.log | head -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r2_bench_beit_n$N.json 2> gpurun_out/r2_bench_beit_n$N.err; tail -c 1500 gpurun_out/r2_bench_beit_n$N.json; tail -3 gpurun_out/r2_bench_beit_n$N.err
timeout 600 python -m pytest tests/test_dist_gpu.py -q -p no:cacheprovider 2>&1 | tail -3
