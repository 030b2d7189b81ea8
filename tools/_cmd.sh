mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_mapper.py -q -m gpu -k "grouped or step_host or random_draws or slabs" > gpurun_out/r02e_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r02e_memcheck.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/probe_bwd.py 7 32 > gpurun_out/r02e_racecheck_backward.log 2>&1; echo "racecheck bwd rc=$?"; tail -2 gpurun_out/r02e_racecheck_backward.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/probe_fwd.py 7 32 > gpurun_out/r02e_racecheck_forward.log 2>&1; echo "racecheck fwd rc=$?"; tail -2 gpurun_out/r02e_racecheck_forward.log
