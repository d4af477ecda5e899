o=gpurun_out/sweep_coalesce_r02d.txt; rm -f $o
python tools/sweep_lanes.py 17 4 16 msm_coalesce=4 msm_tree_rounds=5 >> $o 2>&1
python tools/sweep_lanes.py 17 4 17 msm_coalesce=4 >> $o 2>&1
python tools/sweep_lanes.py 17 4 15 msm_coalesce=4 tree_k=8 >> $o 2>&1
python tools/sweep_lanes.py 18 4 16 msm_coalesce=4 >> $o 2>&1
python tools/sweep_lanes.py 18 4 17 msm_coalesce=4 >> $o 2>&1
python tools/sweep_lanes.py 18 4 16 msm_coalesce=4 msm_tree_rounds=6 >> $o 2>&1
python tools/sweep_lanes.py 20 4 17 msm_coalesce=2 msm_tree_rounds=6 >> $o 2>&1
python tools/sweep_lanes.py 20 4 17 msm_coalesce=2 acc_split=0 >> $o 2>&1
python tools/sweep_lanes.py 20 4 17 msm_coalesce=2 reduce_lk=2 >> $o 2>&1
python tools/sweep_lanes.py 20 3 17 msm_coalesce=4 >> $o 2>&1
cat $o
