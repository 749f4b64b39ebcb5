# ncu evidence for profiles/, part B: the 1-CTA GEMM / conv kernels, the small bandwidth kernels, resizes and normalise
mkdir -p gpurun_out
P="ncu --clock-control none --profile-from-start off"
F="$P --set full --import-source off -f"
$F -k "regex:gemm_tcgen05_kernel|gemm_tcgen05_persist" -c 3 -o gpurun_out/r02_gemm1 python tools/profile_step.py depth_beit512 > gpurun_out/prof_c.log 2>&1
$F -k "regex:layernorm|preprocess|assemble|concat_readout|attention_cls_row|im2col" -s 4 -c 6 -o gpurun_out/r02_small python tools/profile_step.py depth_beit512 > gpurun_out/prof_d.log 2>&1
$F -k "regex:resize|minmax_f32|quantize" -c 5 -o gpurun_out/r02_resize python tools/profile_step.py depth_beit512 > gpurun_out/prof_e.log 2>&1
$F -k "regex:clb_final|attractor|resize_add|select_softplus|tta_combine|zoe_preprocess|attention_small|layernorm_post" -c 12 -o gpurun_out/r02_zoe python tools/profile_step.py zoedepth_nk768 4 > gpurun_out/prof_h.log 2>&1
du -sh gpurun_out; ls -la gpurun_out/*.ncu-rep
