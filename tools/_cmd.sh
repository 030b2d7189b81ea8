mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r3r_tests.log 2>&1; tail -3 gpurun_out/r3r_tests.log
for sp in 1 0; do echo "== NRT_SPLIT_ADAM=$sp"
NRT_SPLIT_ADAM=$sp python tools/probe_cfg2.py 2048 32 2>&1 | head -1
NRT_SPLIT_ADAM=$sp python tools/probe_cfg2.py 4096 117 office0 2>&1 | head -1
done > gpurun_out/r3r_probe.log 2>&1; cat gpurun_out/r3r_probe.log
