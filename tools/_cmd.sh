mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r02f_gpu_tests.log 2>&1; tail -2 gpurun_out/r02f_gpu_tests.log
python bench.py > gpurun_out/r02f_bench_1gpu.json 2> gpurun_out/r02f_bench_1gpu.err; tail -c 300 gpurun_out/r02f_bench_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02f_bench_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['losses_finite'])
print(d['kernels']); print(d['roofline']['frac'], d['roofline']['hash_gather'], d['sweep']['ms'])
for k,v in d['configs'].items(): print(k, v['ms_per_step'], v.get('fwd_ms'))
print(d['e2e_dropin']['ms_per_step'], d['e2e_mapper']['ms_per_call'], d['clocks'])
PY
