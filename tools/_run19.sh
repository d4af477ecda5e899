python -m pytest tests -x -q -m gpu > gpurun_out/pytest_r02n.log 2>&1; tail -2 gpurun_out/pytest_r02n.log
python bench.py > gpurun_out/bench_r02n_n1.json 2> gpurun_out/bench_r02n_n1.err; tail -2 gpurun_out/bench_r02n_n1.err
python bench.py --impl reference > gpurun_out/bench_r02n_ref.json 2>> gpurun_out/bench_r02n_n1.err
ncu --set full --clock-control none --import-source on -k regex:k_r1_add -s 1 -c 1 -o gpurun_out/prof_r1_add_r02n -f python tools/prof_blob.py 2048 2 15 > gpurun_out/ncu_full2.log 2>&1
ncu -i gpurun_out/prof_r1_add_r02n.ncu-rep --page raw --csv > gpurun_out/prof_r1_add_r02n_raw.csv 2>/dev/null
tail -2 gpurun_out/ncu_full2.log
