ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_l2_c20.csv python tools/prof_msm_fused.py 20 2 2 msm_window_bits=20 > gpurun_out/ncu_l2.log 2>&1
tail -2 gpurun_out/ncu_l2.log
