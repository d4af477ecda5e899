ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fused_n20b.csv python tools/prof_msm_fused.py 20 2 2 > gpurun_out/ncu_fused20b.log 2>&1
tail -1 gpurun_out/ncu_fused20b.log
