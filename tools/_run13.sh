ncu --set full --clock-control none --import-source on -k regex:k_seg_add_lean -s 5 -c 1 -o gpurun_out/prof_seg_add_lean_r02 -f python tools/prof_msm.py 20 2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/prof_seg_add_lean_r02.ncu-rep --page raw --csv > gpurun_out/prof_seg_add_lean_r02_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r02l_bench.csv python bench.py --steps 8 --warmup 3 --blobs 0 > gpurun_out/ncu_bench.log 2>&1
tail -c 300 gpurun_out/ncu_bench.log
ls -la gpurun_out/prof_seg_add_lean_r02.ncu-rep
