set -x
timeout 900 python -m pytest tests/test_gpu_hnsw.py -x -q 2>&1 | tail -15
timeout 900 python scripts/hnsw_bench.py --rows 50000 > gpurun_out/hnsw_bench_50k_v4.json 2> gpurun_out/hnsw_bench_50k_v4.err; tail -3 gpurun_out/hnsw_bench_50k_v4.err; cat gpurun_out/hnsw_bench_50k_v4.json
