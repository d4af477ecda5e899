o=gpurun_out/sweep_window_r02c.txt; rm -f $o
for c in 15 16 17; do timeout 300 python tools/sweep_lanes.py 17 4 $c msm_coalesce=4 >> $o 2>&1; done
for c in 16 17; do timeout 300 python tools/sweep_lanes.py 18 4 $c msm_coalesce=4 >> $o 2>&1; done
for c in 16 17; do timeout 300 python tools/sweep_lanes.py 19 4 $c msm_coalesce=4 >> $o 2>&1; done
timeout 300 python tools/sweep_lanes.py 20 3 17 msm_coalesce=4 >> $o 2>&1
timeout 300 python tools/sweep_lanes.py 20 4 17 msm_coalesce=2 msm_tree_rounds=6 >> $o 2>&1
cat $o
