mkdir -p gpurun_out
python -m pytest tests/test_mapper.py -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 50 --warmup 10 --no-side-configs --no-torch-gpu-baseline --no-cpu-baseline --sweep-rays 0 --no-dropin > gpurun_out/r3s_bench.json 2> gpurun_out/r3s_bench.err; tail -c 400 gpurun_out/r3s_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r3s_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['losses_finite'])"
