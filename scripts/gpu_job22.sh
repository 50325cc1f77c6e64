set -x
timeout 900 python -m pytest tests/test_gpu_tensor.py -x -q 2>&1 | tail -6
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pair.json 2> gpurun_out/bench_pair.err; tail -2 gpurun_out/bench_pair.err; cat gpurun_out/bench_pair.json
timeout 900 python bench.py --workload flat_int8_cos_50M_d512_k10_b4096 --rows 6250000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_int8_shard2.json 2> gpurun_out/bench_int8_shard2.err; tail -2 gpurun_out/bench_int8_shard2.err; cat gpurun_out/bench_int8_shard2.json
timeout 900 python bench.py --workload flat_bf16_ip_20M_d1024_k100_b1024 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16_cfg4_pair.json 2> /dev/null; cat gpurun_out/bench_bf16_cfg4_pair.json
