o=gpurun_out/sweep_l2_r02.txt; rm -f $o
timeout 300 python tools/sweep_lanes.py 15 1 12 msm_l2=1 >> $o 2>&1
timeout 300 python tools/sweep_lanes.py 17 2 17 msm_l2=1 msm_coalesce=2 >> $o 2>&1
timeout 300 python tools/sweep_lanes.py 20 4 17 msm_l2=1 msm_coalesce=2 >> $o 2>&1
timeout 300 python tools/sweep_lanes.py 20 1,4 19 >> $o 2>&1
timeout 300 python tools/sweep_lanes.py 20 4 19 msm_coalesce=2 >> $o 2>&1
timeout 300 python tools/sweep_lanes.py 20 1,4 20 >> $o 2>&1
timeout 300 python tools/sweep_lanes.py 20 4 20 msm_coalesce=2 >> $o 2>&1
tail -c 3000 $o
