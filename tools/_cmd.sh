mkdir -p gpurun_out
N=$1
if [ "$N" = "2" ]; then
python -m pytest tests/test_parallel_gpu.py -q -m gpu > gpurun_out/r02e_dp_2gpu_test.log 2>&1; tail -2 gpurun_out/r02e_dp_2gpu_test.log
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 50 --warmup 10 > gpurun_out/r02e_bench_${N}gpu.json 2> gpurun_out/r02e_bench_${N}gpu.err
tail -c 200 gpurun_out/r02e_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02e_bench_${N}gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['parallelism'])
for k,v in (d.get('configs') or {}).items(): print(k, v['ms_per_step'], v['rays_per_s'])
PY
