mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zoe_gpu.py -q -s -p no:cacheprovider > gpurun_out/r2_zoe.log 2>&1; echo "rc=$?" >> gpurun_out/r2_zoe.log
grep -E "^\[precision\]|passed|failed|^FAILED|Error|rc=" gpurun_out/r2_zoe.log | tail -30
timeout 900 python -m pytest tests/test_models_baseline_gpu.py -q -s -p no:cacheprovider -k batch32 > gpurun_out/r2_b32.log 2>&1
grep -E "^\[precision\]|passed|failed|^FAILED|rc=" gpurun_out/r2_b32.log | tail
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd4 -c 1 -o gpurun_out/prof_r2_attn4 -f python tools/bench_attention.py beit > gpurun_out/r2_ncu_attn4.log 2>&1; tail -3 gpurun_out/r2_ncu_attn4.log
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_vit_ops_gpu.py -q -x -k "test_attention and (257 or 16-16 or 8-6 or 700)" -p no:cacheprovider > gpurun_out/r2_memcheck_attn4.log 2>&1; tail -4 gpurun_out/r2_memcheck_attn4.log
timeout 500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_vit_ops_gpu.py -q -x -k "test_attention and (257-2 or 8-6)" -p no:cacheprovider > gpurun_out/r2_racecheck_attn4.log 2>&1; tail -4 gpurun_out/r2_racecheck_attn4.log
