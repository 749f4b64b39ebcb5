# ncu evidence, part C: the kernels written or rewritten late in round 2 (ZoeDepth head v2, BOOST merge net + glue) and the final attention build
mkdir -p gpurun_out
P="ncu --clock-control none --profile-from-start off"
F="$P --set full --import-source off -f"
timeout 400 $F -k "regex:clb_final|attractor_kernel|attention_small" -c 6 -o gpurun_out/r02_zoe_head_v2 python tools/profile_step.py zoedepth_nk768 8 > gpurun_out/prof_h.log 2>&1
timeout 400 $F -k "regex:unet_|sum_chunks" -c 14 -o gpurun_out/r02_unet python tools/profile_unet.py unet > gpurun_out/prof_i.log 2>&1
timeout 400 $F -k "regex:boost_|leres_stem" -c 14 -o gpurun_out/r02_boost_glue python tools/profile_unet.py patch > gpurun_out/prof_j.log 2>&1
timeout 400 $F -k regex:attention_fwd4 -s 2 -c 1 -o gpurun_out/r02_attn_final python tools/profile_step.py depth_beit512 > gpurun_out/prof_k.log 2>&1
du -sh gpurun_out; ls -la gpurun_out/r02_zoe_head_v2.ncu-rep gpurun_out/r02_unet.ncu-rep gpurun_out/r02_boost_glue.ncu-rep gpurun_out/r02_attn_final.ncu-rep; tail -2 gpurun_out/prof_h.log gpurun_out/prof_i.log gpurun_out/prof_j.log gpurun_out/prof_k.log 2>&1 | tail -12
