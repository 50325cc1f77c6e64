set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 3 --warmup 3 --mode 1 --batch 16 --no-cpu-baseline > gpurun_out/bench_exact_b16.json 2> gpurun_out/bench_exact_b16.err; tail -3 gpurun_out/bench_exact_b16.err; cat gpurun_out/bench_exact_b16.json
VSGPU_LEGACY_SCAN=1 timeout 600 python bench.py --steps 3 --warmup 3 --mode 1 --batch 16 --no-cpu-baseline > gpurun_out/bench_exact_b16_legacy.json 2> gpurun_out/bench_exact_b16_legacy.err; cat gpurun_out/bench_exact_b16_legacy.json
timeout 600 python bench.py --steps 3 --warmup 3 --mode 1 --batch 1 --no-cpu-baseline > gpurun_out/bench_exact_b1.json 2> gpurun_out/bench_exact_b1.err; cat gpurun_out/bench_exact_b1.json
timeout 600 python bench.py --steps 3 --warmup 3 --mode 1 --batch 8 --no-cpu-baseline > gpurun_out/bench_exact_b8.json 2> gpurun_out/bench_exact_b8.err; cat gpurun_out/bench_exact_b8.json
timeout 900 python scripts/hnsw_bench.py --rows 50000 > gpurun_out/hnsw_bench_50k_v2.json 2> gpurun_out/hnsw_bench_50k_v2.err; tail -3 gpurun_out/hnsw_bench_50k_v2.err; cat gpurun_out/hnsw_bench_50k_v2.json
