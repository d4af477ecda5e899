o=gpurun_out/sweep_coalesce_r02c.txt; rm -f $o
python tools/sweep_lanes.py 20 4 16 msm_coalesce=2 >> $o 2>&1
python tools/sweep_lanes.py 20 4 17 msm_coalesce=2 >> $o 2>&1
python tools/sweep_lanes.py 20 4 18 msm_coalesce=2 >> $o 2>&1
python tools/sweep_lanes.py 20 4 17 msm_coalesce=2 msm_tree_rounds=4 >> $o 2>&1
python tools/sweep_lanes.py 20 4 17 >> $o 2>&1
python tools/sweep_lanes.py 19 4 17 msm_coalesce=2 >> $o 2>&1
python tools/sweep_lanes.py 19 4 16 msm_coalesce=2 >> $o 2>&1
python tools/sweep_lanes.py 19 4 16 msm_coalesce=4 >> $o 2>&1
cat $o
