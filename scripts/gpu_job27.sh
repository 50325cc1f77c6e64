set -x
timeout 600 python -m pytest tests/test_gpu_tensor.py -x -q 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_hnsw.py -x -q 2>&1 | tail -4
VSGPU_HNSW_PROFILE=1 timeout 900 python scripts/hnsw_bench.py --rows 100000 > gpurun_out/hnsw_bench_100k_batch2.json 2> gpurun_out/hnsw_bench_100k_batch2.err; grep insert gpurun_out/hnsw_bench_100k_batch2.err | tail -2 | cut -c1-200; cat gpurun_out/hnsw_bench_100k_batch2.json
