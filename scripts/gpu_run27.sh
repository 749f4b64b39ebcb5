mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_boost_gpu.py -m gpu -q -s -p no:cacheprovider -k "unet or glue or estimateboost" > gpurun_out/r2_pytest27.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest27.log
grep -E "precision|passed|failed|^FAILED|^ERROR|rc=|Error|capture failed" gpurun_out/r2_pytest27.log | tail -12
timeout 600 python tools/bench_boost_parts.py > gpurun_out/r02_boost_parts.txt 2>&1; cat gpurun_out/r02_boost_parts.txt | tail -12
