python -m pytest tests/test_gpu_coalesce.py tests/test_gpu_parity.py -x -q -m gpu -k "coalesce or fused or msm or commit or c3" > gpurun_out/pytest_r02m.log 2>&1; tail -2 gpurun_out/pytest_r02m.log
o=gpurun_out/sweep_r02m.txt; rm -f $o
python tools/sweep_lanes.py 20 1 >> $o 2>&1
python tools/sweep_lanes.py 20 4 msm_coalesce=2 >> $o 2>&1
python tools/sweep_lanes.py 17 4 msm_coalesce=4 >> $o 2>&1
python tools/sweep_lanes.py 17 1 >> $o 2>&1
cat $o
