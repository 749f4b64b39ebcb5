mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_leres_gpu.py -q -s -p no:cacheprovider -x > gpurun_out/r2_run9_leres.log 2>&1; echo "rc=$?" >> gpurun_out/r2_run9_leres.log
grep -E "^\[precision\]|passed|failed|^FAILED|rc=|^E  |Error" gpurun_out/r2_run9_leres.log | tail -30
