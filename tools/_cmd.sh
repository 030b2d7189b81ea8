# Evidence run on one B200 (what produced profiles/r02e_*):  gpurun --timeout 1500 -- 'bash tools/_cmd.sh'
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02e_smoke.log 2>&1; tail -2 gpurun_out/r02e_smoke.log
python -m pytest tests -q -m gpu > gpurun_out/r02e_gpu_tests.log 2>&1; tail -3 gpurun_out/r02e_gpu_tests.log
python bench.py > gpurun_out/r02e_bench_1gpu.json 2> gpurun_out/r02e_bench_1gpu.err; tail -c 300 gpurun_out/r02e_bench_1gpu.err
python bench.py --impl reference > gpurun_out/r02e_bench_reference_arm.json 2> gpurun_out/r02e_bench_reference_arm.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02e_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-torch-gpu-baseline --no-side-configs --no-dropin --sweep-rays 0 > gpurun_out/r02e_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:render_fwd_ws_kernel|decode_bwd_q_kernel' -s 6 -c 2 -o gpurun_out/r02e_full -f python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-torch-gpu-baseline --no-side-configs --no-dropin --sweep-rays 0 > gpurun_out/r02e_full.log 2>&1
# multi-GPU (gpurun --gpus N):  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus N --steps 50 --warmup 10
# stage accounting:             python -m torch.distributed.run ... tools/probe_dp.py
