mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --deselect tests/test_models_baseline_gpu.py -p no:cacheprovider > gpurun_out/r2_pytest_a.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_a.log
python -m pytest tests/test_models_baseline_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_pytest_b.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_b.log
python tests/diag_precision.py vits 140 > gpurun_out/r2_diag_vits.log 2>&1
python tests/diag_precision.py vitl 518 > gpurun_out/r2_diag_vitl.log 2>&1
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "golden and smooth and (polylines_sharp or naive_interpolating)" -p no:cacheprovider > gpurun_out/r2_memcheck_stereo.log 2>&1
timeout 400 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "golden and smooth and polylines_sharp" -p no:cacheprovider > gpurun_out/r2_racecheck_stereo.log 2>&1
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_vit_ops_gpu.py -q -x -k "test_attention and (257 or 16-16 or 8-6)" -p no:cacheprovider > gpurun_out/r2_memcheck_attn.log 2>&1
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gemm_gpu.py -q -x -k "4100 or 128-128-64 or 19-19" -p no:cacheprovider > gpurun_out/r2_memcheck_gemm.log 2>&1
tail -3 gpurun_out/r2_pytest_a.log gpurun_out/r2_pytest_b.log
