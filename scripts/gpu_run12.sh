mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd4 -c 1 -o gpurun_out/prof_r2_attn4e_dav2 -f python tools/bench_attention.py dav2 > gpurun_out/r2_ncu_attn4e.log 2>&1; tail -2 gpurun_out/r2_ncu_attn4e.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd4 -c 1 -o gpurun_out/prof_r2_attn4e_beit -f python tools/bench_attention.py beit >> gpurun_out/r2_ncu_attn4e.log 2>&1; tail -2 gpurun_out/r2_ncu_attn4e.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sub --no-funnel > gpurun_out/r2_bench_default_c.json 2> gpurun_out/r2_bench_default_c.err; tail -c 700 gpurun_out/r2_bench_default_c.json
