mkdir -p gpurun_out
timeout 120 tools/micro/random_gather > gpurun_out/r3t_random_gather.log 2>&1
cat gpurun_out/r3t_random_gather.log
