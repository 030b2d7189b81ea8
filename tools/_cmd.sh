timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2i_tests.log 2>&1; tail -5 gpurun_out/r2i_tests.log
P="python tools/probe_bwd.py"
PROBE_WARPS=1 NRT_BWD_DEBUG=8 $P 4096 117 > gpurun_out/r2i_probe.log 2>&1
NRT_BWD_DEBUG=24 $P 4096 117 >> gpurun_out/r2i_probe.log 2>&1
NRT_BWD_DEBUG=8 $P 32768 117 >> gpurun_out/r2i_probe.log 2>&1
NRT_BWD_DEBUG=8 $P 2148 32 >> gpurun_out/r2i_probe.log 2>&1
NRT_BWD_IMPL=tc $P 4096 117 >> gpurun_out/r2i_probe.log 2>&1
grep -v "mlp \|scat " gpurun_out/r2i_probe.log; grep " 0 scat\| 8 scat\|15 scat\|16 mlp\|17 mlp" gpurun_out/r2i_probe.log | head -10
