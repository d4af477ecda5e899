python -m pytest tests -x -q -m gpu > gpurun_out/pytest_r02q.log 2>&1; tail -2 gpurun_out/pytest_r02q.log
python bench.py > gpurun_out/bench_r02q_n1.json 2> gpurun_out/bench_r02q_n1.err; tail -2 gpurun_out/bench_r02q_n1.err
python bench.py --impl reference > gpurun_out/bench_r02q_ref.json 2>> gpurun_out/bench_r02q_n1.err
python __graft_entry__.py smoke 2>&1 | tail -2
