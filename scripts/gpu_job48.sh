set -x
timeout 900 python -m pytest tests/test_gpu_flat.py tests/test_gpu_flat_multi.py tests/test_gpu_tiered.py -x -q 2>&1 | tail -4
timeout 600 python scripts/cfg1_latency.py > gpurun_out/cfg1_latency_v3.json 2> gpurun_out/cfg1_latency.err; tail -3 gpurun_out/cfg1_latency.err; cat gpurun_out/cfg1_latency_v3.json
