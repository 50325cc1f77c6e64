set -x
timeout 600 python scripts/cfg1_latency.py > gpurun_out/cfg1_latency.json 2> gpurun_out/cfg1_latency.err; tail -3 gpurun_out/cfg1_latency.err; cat gpurun_out/cfg1_latency.json
