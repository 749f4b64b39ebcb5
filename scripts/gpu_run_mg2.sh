mkdir -p gpurun_out
N=${1:-2}
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/r02_dist_check_n$N.log 2>&1; grep -E "dist_check|NCCL INFO (Using|comm .* Init COMPLETE|Connected all|NVLS)" gpurun_out/r02_dist_check_n$N.log | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r02_bench_depth_beit512_n$N.json 2> gpurun_out/r02_bench_beit_n$N.err; tail -c 700 gpurun_out/r02_bench_depth_beit512_n$N.json; tail -3 gpurun_out/r02_bench_beit_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --workload boost_res101_2048 --steps 5 --warmup 3 > gpurun_out/r02_bench_boost_n$N.json 2> gpurun_out/r02_bench_boost_n$N.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_boost_n$N.json').read().strip().splitlines()[-1])
print('boost n=$N', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['parallelism'])
PY
tail -3 gpurun_out/r02_bench_boost_n$N.err
