for impl in peer; do
NRT_DP_IMPL=$impl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/r2m_bench8_$impl.json 2> gpurun_out/r2m_bench8_$impl.err; echo rc=$?
grep -v "^\[W\|^$\|\*\*\*\|Warning\|symm.enable" gpurun_out/r2m_bench8_$impl.err | tail -5
python -c "
import json;d=json.load(open('gpurun_out/r2m_bench8_$impl.json'));print('$impl',d['value'],d['ms_per_step'],d['e2e']['ms_per_step'],d['gpu_launches']); print({k:(v['ms_per_step'],v['rays_per_s']) for k,v in d['configs'].items()})"
done
NRT_DP_IMPL=peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 30 --warmup 5 --no-side-configs > gpurun_out/r2m_bench4_peer.json 2> gpurun_out/r2m_bench4_peer.err; echo rc=$?
python -c "
import json;d=json.load(open('gpurun_out/r2m_bench4_peer.json'));print('peer4',d['value'],d['ms_per_step'],d['e2e']['ms_per_step'])"
