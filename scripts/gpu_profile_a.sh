# ncu evidence for profiles/, part A: launch lists of one eager step per workload + `--set full` captures of attention, the 2-SM GEMM family and the stereo / normal-map kernels
mkdir -p gpurun_out
P="ncu --clock-control none --profile-from-start off"
$P --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_depth_beit512.csv python tools/profile_step.py depth_beit512 > gpurun_out/prof_launch_beit.log 2>&1
$P --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_dav2_stereo.csv python tools/profile_step.py dav2_stereo 16 > gpurun_out/prof_launch_dav2.log 2>&1
$P --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_zoedepth_nk768.csv python tools/profile_step.py zoedepth_nk768 8 > gpurun_out/prof_launch_zoe.log 2>&1
F="$P --set full --import-source off -f"
$F -k regex:attention_fwd4 -s 2 -c 1 -o gpurun_out/r02_attn python tools/profile_step.py depth_beit512 > gpurun_out/prof_a.log 2>&1
$F -k regex:attention_fwd4 -s 2 -c 1 -o gpurun_out/r02_attn_dav2 python tools/profile_step.py dav2_stereo 16 > gpurun_out/prof_g.log 2>&1
$F -k regex:gemm_tcgen05_2sm_kernel -s 8 -c 4 -o gpurun_out/r02_gemm2sm python tools/profile_step.py depth_beit512 > gpurun_out/prof_b.log 2>&1
$F -k "regex:stereo_row|normalmap|minmax_u16" -c 3 -o gpurun_out/r02_stereo python tools/profile_step.py stereo2048 > gpurun_out/prof_f.log 2>&1
du -sh gpurun_out; ls -la gpurun_out/*.ncu-rep
