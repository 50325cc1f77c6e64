set -x
timeout 900 python bench.py --workload flat_fp32_l2_10M_d768_k100_b1024 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_l2_10M.json 2> gpurun_out/bench_l2_10M.err; tail -3 gpurun_out/bench_l2_10M.err; cat gpurun_out/bench_l2_10M.json
