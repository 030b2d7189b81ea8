mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r3d_tests.log 2>&1; tail -3 gpurun_out/r3d_tests.log
python bench.py > gpurun_out/r3d_bench.json 2> gpurun_out/r3d_bench.err; tail -c 600 gpurun_out/r3d_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3d_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches'])
print(d['kernels']); print(d['sweep']['ms'], d['sweep']['frac_of_hbm_peak'])
for k,v in d['configs'].items(): print(k, v['ms_per_step'], v.get('fwd_ms'))
print(d.get('e2e_dropin'))
PY
