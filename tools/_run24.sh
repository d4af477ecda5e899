python -m pytest tests -x -q -m gpu > gpurun_out/pytest_r02p.log 2>&1; tail -2 gpurun_out/pytest_r02p.log
python tools/prof_blob.py 4096 3 15 > gpurun_out/blob_passthrough.txt 2>&1; cat gpurun_out/blob_passthrough.txt
