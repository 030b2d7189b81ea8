timeout 600 python -m pytest tests/test_parallel_gpu.py -x -q -m gpu -s > gpurun_out/r2v_dp_test.log 2>&1; tail -12 gpurun_out/r2v_dp_test.log
timeout 300 python -m pytest tests/test_scale_properties.py tests/test_mapper.py -x -q -m gpu > gpurun_out/r2v_tests.log 2>&1; tail -2 gpurun_out/r2v_tests.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-torch-gpu-baseline --no-dropin --no-side-configs --sweep-rays 0 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read());print(d['ms_per_step'],d['kernels'])"
