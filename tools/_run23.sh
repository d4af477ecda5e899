python -m pytest tests/test_gpu_coalesce.py tests/test_gpu_parity.py -x -q -m gpu -k "coalesce or fused or msm or commit or c3" > gpurun_out/pytest_r02o.log 2>&1; tail -2 gpurun_out/pytest_r02o.log
o=gpurun_out/sweep_l2_r02b.txt; rm -f $o
timeout 300 python tools/sweep_lanes.py 15 1 12 msm_l2=1 >> $o 2>&1
timeout 300 python tools/sweep_lanes.py 20 1,4 17 >> $o 2>&1
timeout 300 python tools/sweep_lanes.py 20 4 17 msm_coalesce=2 >> $o 2>&1
timeout 300 python tools/sweep_lanes.py 20 4 19 msm_coalesce=2 >> $o 2>&1
timeout 300 python tools/sweep_lanes.py 20 1,4 20 >> $o 2>&1
timeout 300 python tools/sweep_lanes.py 20 4 20 msm_coalesce=2 >> $o 2>&1
timeout 300 python tools/sweep_lanes.py 17 4 15 msm_coalesce=4 >> $o 2>&1
timeout 300 python tools/sweep_lanes.py 17 1 >> $o 2>&1
tail -c 3000 $o
