timeout 600 python -m pytest tests/test_scale_properties.py tests/test_cuda_parity.py tests/test_mapper.py tests/test_coslam_mapper.py -x -q -m gpu > gpurun_out/r2t_tests.log 2>&1; tail -3 gpurun_out/r2t_tests.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-torch-gpu-baseline --no-side-configs --no-dropin --sweep-rays 0 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2t_bench.json'));print(d['value'],d['ms_per_step'],d['kernels'],d['roofline']['frac'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2t_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-torch-gpu-baseline --no-side-configs --no-dropin --sweep-rays 0 > /dev/null 2>&1
python tools/ncu_summary.py launches gpurun_out/r2t_launches.csv | head -12
