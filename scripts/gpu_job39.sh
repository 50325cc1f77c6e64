set -x
timeout 900 python bench.py --workload flat_bf16_ip_20M_d1024_k100_b1024 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16_cfg4_v4.json 2> gpurun_out/bench_bf16_cfg4_v4.err; tail -2 gpurun_out/bench_bf16_cfg4_v4.err; cat gpurun_out/bench_bf16_cfg4_v4.json
timeout 900 python scripts/hnsw_bench.py --rows 100000 > gpurun_out/hnsw_bench_100k_v4.json 2> gpurun_out/hnsw_bench_100k_v4.err; cat gpurun_out/hnsw_bench_100k_v4.json
