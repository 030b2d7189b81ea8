timeout 400 python -m pytest tests/test_scale_properties.py tests/test_cuda_parity.py tests/test_mapper.py -x -q -m gpu > gpurun_out/r2r_tests.log 2>&1; tail -3 gpurun_out/r2r_tests.log
P="python tools/probe_bwd.py"
PROBE_WARPS=1 NRT_BWD_DEBUG=8 timeout 100 $P 4096 117 > gpurun_out/r2r_probe.log 2>&1
NRT_BWD_DEBUG=8 timeout 100 $P 32768 117 >> gpurun_out/r2r_probe.log 2>&1
NRT_BWD_DEBUG=8 timeout 100 $P 2148 32 >> gpurun_out/r2r_probe.log 2>&1
NRT_BWD_IMPL=tc timeout 100 $P 4096 117 >> gpurun_out/r2r_probe.log 2>&1
grep -v "mlp \|scat " gpurun_out/r2r_probe.log | tail -32; grep " 0 scat\| 8 scat\|16 mlp\|17 mlp" gpurun_out/r2r_probe.log | head -4
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-torch-gpu-baseline --no-side-configs --no-dropin > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2r_bench.json'));print(d['value'],d['ms_per_step'],d['kernels'],d['roofline']['frac'],d['sweep']['ms'],d['sweep']['frac_of_hbm_peak'])"
