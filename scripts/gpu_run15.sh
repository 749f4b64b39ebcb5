mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_pytest_final.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_final.log
grep -E "passed|failed|^FAILED|^ERROR|rc=" gpurun_out/r2_pytest_final.log | tail -10
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
