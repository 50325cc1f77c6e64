set -x
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6
timeout 600 python scripts/cfg1_latency.py > gpurun_out/cfg1_latency_v2.json 2> gpurun_out/cfg1_latency.err; tail -3 gpurun_out/cfg1_latency.err; cat gpurun_out/cfg1_latency_v2.json
