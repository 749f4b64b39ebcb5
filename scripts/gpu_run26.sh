mkdir -p gpurun_out
P="ncu --clock-control none --profile-from-start off"
timeout 600 $P --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_unet.csv python tools/profile_unet.py unet > gpurun_out/prof_unet.log 2>&1
timeout 600 $P --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_leres896.csv python tools/profile_unet.py leres > gpurun_out/prof_leres.log 2>&1
tail -2 gpurun_out/prof_unet.log gpurun_out/prof_leres.log; wc -l gpurun_out/r02_launches_unet.csv gpurun_out/r02_launches_leres896.csv
