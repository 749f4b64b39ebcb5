# ncu evidence, part D: the rewritten log-binomial kernel and the merge net's direct end kernels after the bank-conflict fix
mkdir -p gpurun_out
P="ncu --clock-control none --profile-from-start off"
F="$P --set full --import-source off -f"
timeout 300 $F -k "regex:clb_final" -c 1 -o gpurun_out/r02_zoe_clb_v2 python tools/profile_step.py zoedepth_nk768 8 > gpurun_out/prof_l.log 2>&1
timeout 300 $F -k "regex:unet_first_kernel|unet_last_kernel|unet_up_cols|unet_interleave" -c 6 -o gpurun_out/r02_unet_ends python tools/profile_unet.py unet > gpurun_out/prof_m.log 2>&1
ls -la gpurun_out/r02_zoe_clb_v2.ncu-rep gpurun_out/r02_unet_ends.ncu-rep
