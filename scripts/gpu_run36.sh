mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_boost_gpu.py -m gpu -q -s -p no:cacheprovider -k "unet or glue or estimateboost" > gpurun_out/r2_pytest36.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest36.log
grep -E "precision\] (boost|merge)|passed|failed|^FAILED|^ERROR|rc=" gpurun_out/r2_pytest36.log | tail -8
timeout 600 python tools/bench_boost_parts.py > gpurun_out/r02_boost_parts_v3.txt 2>&1; grep -E "U-Net|one patch" gpurun_out/r02_boost_parts_v3.txt
