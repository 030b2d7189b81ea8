timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2j_tests.log 2>&1; tail -8 gpurun_out/r2j_tests.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo rc=$?
tail -3 gpurun_out/r2j_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['kernels'], d['roofline']['frac'])
print('sweep', d.get('sweep')); print('dropin', d.get('e2e_dropin')); print('torch', d.get('torch_gpu_baseline')); print('cpu', d.get('cpu_baseline'))
print('configs', json.dumps(d.get('configs'), indent=1))
PY
