timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2x_tests.log 2>&1; tail -3 gpurun_out/r2x_tests.log
for i in 1 2; do timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-torch-gpu-baseline --no-dropin --sweep-rays 0 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['ms_per_step'],d['kernels'],{k:(v['ms_per_step'],v.get('fwd_ms')) for k,v in d['configs'].items()})"; done
NRT_FWD_DEBUG=1 python tools/probe_fwd.py 4096 117 2>&1 | head -7
NRT_FWD_DEBUG=1 python tools/probe_fwd.py 2148 32 2>&1 | head -7
