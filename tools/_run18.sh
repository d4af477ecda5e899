ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r02m_blob_4096.csv python tools/prof_blob.py 4096 1 15 > gpurun_out/ncu_blob.log 2>&1
tail -2 gpurun_out/ncu_blob.log
