set -x
timeout 900 python bench.py --workload flat_int8_cos_50M_d512_k10_b4096 --rows 6250000 --steps 5 --warmup 3 --cpu-seconds 8 > gpurun_out/bench_int8_shard.json 2> gpurun_out/bench_int8_shard.err; tail -3 gpurun_out/bench_int8_shard.err; cat gpurun_out/bench_int8_shard.json
timeout 900 python bench.py --workload flat_int8_cos_50M_d512_k10_b4096 --rows 6250000 --steps 2 --warmup 1 --mode 1 --batch 256 --no-cpu-baseline > gpurun_out/bench_int8_shard_exact.json 2> gpurun_out/bench_int8_shard_exact.err; cat gpurun_out/bench_int8_shard_exact.json
timeout 900 python bench.py --workload flat_bf16_ip_20M_d1024_k100_b1024 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16_cfg4.json 2> gpurun_out/bench_bf16_cfg4.err; tail -2 gpurun_out/bench_bf16_cfg4.err; cat gpurun_out/bench_bf16_cfg4.json
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
