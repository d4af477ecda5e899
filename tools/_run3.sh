python -m pytest tests/test_gpu_coalesce.py tests/test_gpu_parity.py -x -q -m gpu -k "coalesce or fused or msm or commit" > gpurun_out/pytest_coalesce_b.log 2>&1
tail -4 gpurun_out/pytest_coalesce_b.log
o=gpurun_out/sweep_coalesce_r02b.txt; rm -f $o
python tools/sweep_lanes.py 17 4 >> $o 2>&1
python tools/sweep_lanes.py 17 2,4 msm_coalesce=4 >> $o 2>&1
python tools/sweep_lanes.py 17 4 msm_coalesce=4 acc_split=0 >> $o 2>&1
python tools/sweep_lanes.py 17 4 msm_coalesce=4 reduce_lk=3 >> $o 2>&1
python tools/sweep_lanes.py 17 4 msm_coalesce=4 reduce_lk=0 >> $o 2>&1
python tools/sweep_lanes.py 17 4 msm_coalesce=4 msm_tree_rounds=5 >> $o 2>&1
python tools/sweep_lanes.py 17 4 16 msm_coalesce=4 >> $o 2>&1
python tools/sweep_lanes.py 20 4 >> $o 2>&1
python tools/sweep_lanes.py 20 4 reduce_lk=0 >> $o 2>&1
python tools/sweep_lanes.py 20 4 msm_coalesce=2 >> $o 2>&1
python tools/sweep_lanes.py 20 2,3 msm_coalesce=4 >> $o 2>&1
cat $o
