nvidia-smi -L | head -3
python -m pytest tests/test_gpu_multi.py tests/test_gpu_coalesce.py tests/test_gpu_group.py -x -q -m gpu > gpurun_out/pytest_r02k_n2.log 2>&1; tail -3 gpurun_out/pytest_r02k_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_r02k_n2.json 2> gpurun_out/bench_r02k_n2.err; tail -3 gpurun_out/bench_r02k_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02k_n2.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['one_call_at_a_time'], d['config']['calls_in_flight'], d['roofline']['frac'], d['roofline']['canonical_frac'])
print(d.get('single_process'))
print(d['blob_batch']['value'] if 'blob_batch' in d else None)
PY
