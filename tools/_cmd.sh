mkdir -p gpurun_out
for mc in 1 0; do
NRT_DP_MULTICAST=$mc timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/probe_dp.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$\|NCCL"
done > gpurun_out/r3p_dp_stages_8gpu.log
cat gpurun_out/r3p_dp_stages_8gpu.log
