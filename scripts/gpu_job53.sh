set -x
timeout 900 python -m pytest tests/test_gpu_hnsw.py tests/test_gpu_tiered.py tests/test_gpu_hnsw_file.py -x -q 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -k "hnsw or cfg5 or config4 or graph" 2>&1 | tail -3
timeout 600 python scripts/hnsw_bench.py --rows 20000 > gpurun_out/hnsw_bench_20k_v5.json 2> gpurun_out/hnsw_bench_20k_v5.err; cat gpurun_out/hnsw_bench_20k_v5.json
