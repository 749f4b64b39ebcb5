mkdir -p gpurun_out
timeout 600 python tools/bench_boost_parts.py > gpurun_out/r02_boost_parts.txt 2>&1; cat gpurun_out/r02_boost_parts.txt | tail -14
