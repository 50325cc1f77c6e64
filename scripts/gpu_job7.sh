set -x
timeout 600 python -m pytest tests/test_gpu_hnsw.py -x -q 2>&1 | tail -5
timeout 900 python scripts/hnsw_bench.py --rows 50000 --ref > gpurun_out/hnsw_bench_50k.json 2> gpurun_out/hnsw_bench_50k.err; tail -3 gpurun_out/hnsw_bench_50k.err; cat gpurun_out/hnsw_bench_50k.json
