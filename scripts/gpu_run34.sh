mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02_pytest_final.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_final.log
grep -E "passed|failed|^FAILED|^ERROR|rc=" gpurun_out/r02_pytest_final.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
