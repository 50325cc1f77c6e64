set -x
which compute-sanitizer
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tensor.py -x -q -k "tensor_path_equals_oracle and 0-1-128-40000" > gpurun_out/san_memcheck_tensor.log 2>&1; echo "memcheck tensor rc=$?"; tail -5 gpurun_out/san_memcheck_tensor.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/cfg1_latency.py > gpurun_out/san_memcheck_small.log 2>&1; echo "memcheck small-select rc=$?"; tail -3 gpurun_out/san_memcheck_small.log | cut -c1-300
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_tensor.py -x -q -k "tensor_path_equals_oracle and 0-1-128-40000" > gpurun_out/san_racecheck_tensor.log 2>&1; echo "racecheck tensor rc=$?"; tail -5 gpurun_out/san_racecheck_tensor.log
