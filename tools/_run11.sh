python -m pytest tests -x -q -m gpu -rs > gpurun_out/pytest_r02l.log 2>&1; tail -12 gpurun_out/pytest_r02l.log | cut -c1-220
o=gpurun_out/sweep_sort_r02b.txt; rm -f $o
python tools/sweep_lanes.py 20 1,4 >> $o 2>&1
python tools/sweep_lanes.py 20 1,4 msm_sort=1 >> $o 2>&1
python tools/sweep_lanes.py 20 4 msm_coalesce=2 >> $o 2>&1
python tools/sweep_lanes.py 19 1,4 >> $o 2>&1
python tools/sweep_lanes.py 19 1,4 msm_sort=1 >> $o 2>&1
python tools/sweep_lanes.py 17 1 >> $o 2>&1
cat $o
