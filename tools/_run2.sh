python tools/prof_msm_fused.py 17 4 3 > gpurun_out/prof_fused_plain.log 2>&1 || exit 1
cat gpurun_out/prof_fused_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_fused_n17.csv python tools/prof_msm_fused.py 17 4 2 > gpurun_out/ncu_fused.log 2>&1
tail -3 gpurun_out/ncu_fused.log
