mkdir -p gpurun_out
python -m pytest tests/test_sampler_gpu.py tests/test_coslam_mapper.py -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-side-configs --no-torch-gpu-baseline --no-cpu-baseline --sweep-rays 0 > gpurun_out/r3u_bench.json 2> gpurun_out/r3u_bench.err; tail -c 300 gpurun_out/r3u_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r3u_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value']); print(d.get('e2e_mapper')); print(d['roofline']['hash_gather'])"
