set -x
python -m pytest tests/test_gpu_coalesce.py -x -q -m gpu > gpurun_out/pytest_coalesce_a.log 2>&1
tail -5 gpurun_out/pytest_coalesce_a.log
for ln in 17 18; do
python tools/sweep_lanes.py $ln 4 >> gpurun_out/sweep_coalesce_r02.txt 2>&1
python tools/sweep_lanes.py $ln 2,4 msm_coalesce=4 >> gpurun_out/sweep_coalesce_r02.txt 2>&1
python tools/sweep_lanes.py $ln 2 msm_coalesce=2 >> gpurun_out/sweep_coalesce_r02.txt 2>&1
done
python tools/sweep_lanes.py 20 4 >> gpurun_out/sweep_coalesce_r02.txt 2>&1
python tools/sweep_lanes.py 20 2 msm_coalesce=4 >> gpurun_out/sweep_coalesce_r02.txt 2>&1
python tools/sweep_lanes.py 20 2,4 msm_coalesce=2 >> gpurun_out/sweep_coalesce_r02.txt 2>&1
cat gpurun_out/sweep_coalesce_r02.txt
