ncu --set full --clock-control none --import-source on -k regex:k_seg_add_lean -s 4 -c 1 -o gpurun_out/prof_seg_add_lean_r1_r02 -f python tools/prof_msm.py 20 2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/prof_seg_add_lean_r1_r02.ncu-rep --page raw --csv > gpurun_out/prof_seg_add_lean_r1_r02_raw.csv 2>/dev/null
