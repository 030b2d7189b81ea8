mkdir -p gpurun_out
NRT_PROFILE_MAPPER=1 python bench.py --steps 5 --warmup 3 --no-side-configs --no-torch-gpu-baseline --no-cpu-baseline --sweep-rays 0 > gpurun_out/r3g_bench.json 2> gpurun_out/r3g_prof.txt
python -c "
import json
d=json.loads(open('gpurun_out/r3g_bench.json').read().strip().splitlines()[-1]); print(d.get('e2e_mapper'))"
python -m pytest tests/test_coslam_mapper.py -x -q 2>&1 | tail -2
