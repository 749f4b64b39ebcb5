# compute-sanitizer on the round-2 kernels (small shapes): memcheck + racecheck summaries for profiles/
mkdir -p gpurun_out
S="compute-sanitizer --print-limit 10"
timeout 500 $S --tool memcheck python -m pytest tests/test_vit_ops_gpu.py -q -x -k "test_attention and (257 or 16-16 or 8-6 or 700 or 20-32)" -p no:cacheprovider > gpurun_out/r02_memcheck_attention.log 2>&1
timeout 500 $S --tool racecheck python -m pytest tests/test_vit_ops_gpu.py -q -x -k "test_attention and (257-2 or 8-6 or 16-16)" -p no:cacheprovider > gpurun_out/r02_racecheck_attention.log 2>&1
timeout 500 $S --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "golden and smooth" -p no:cacheprovider > gpurun_out/r02_memcheck_stereo_normal.log 2>&1
timeout 500 $S --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -k "golden and smooth and (polylines_sharp or naive_interpolating)" -p no:cacheprovider > gpurun_out/r02_racecheck_stereo.log 2>&1
timeout 600 $S --tool memcheck python -m pytest tests/test_zoe_gpu.py tests/test_leres_gpu.py tests/test_video_gpu.py tests/test_model_cabi_gpu.py -q -x -k "(tiny and 3-hw2) or (leres_vs_oracle and hw0) or (bit_exact and 7-) or (native_dav2 and hw1)" -p no:cacheprovider > gpurun_out/r02_memcheck_zoe_leres_video_model.log 2>&1
timeout 500 $S --tool memcheck python -m pytest tests/test_gemm_gpu.py -q -x -k "4100 or 128-128-64 or 19-19" -p no:cacheprovider > gpurun_out/r02_memcheck_gemm.log 2>&1
for f in gpurun_out/r02_*check*.log; do echo "== $f"; grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" $f | tail -2; done
