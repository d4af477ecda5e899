python -m pytest tests -x -q -m gpu > gpurun_out/pytest_r02k.log 2>&1; tail -3 gpurun_out/pytest_r02k.log
python bench.py > gpurun_out/bench_r02k_n1.json 2> gpurun_out/bench_r02k_n1.err; tail -2 gpurun_out/bench_r02k_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02k_n1.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['one_call_at_a_time'], d['config']['calls_in_flight'], d['roofline']['frac'], d['roofline']['canonical_frac'])
PY
