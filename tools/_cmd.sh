timeout 600 python -m pytest tests/test_scale_properties.py tests/test_cuda_parity.py tests/test_mapper.py tests/test_sampler_gpu.py -x -q -m gpu > gpurun_out/r2w_tests.log 2>&1; tail -3 gpurun_out/r2w_tests.log
for i in 1 2; do timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-torch-gpu-baseline --no-dropin --no-side-configs --sweep-rays 0 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['ms_per_step'],d['kernels'])"; done
NRT_BWD_DEBUG=8 python tools/probe_bwd.py 4096 117 2>&1 | head -8
