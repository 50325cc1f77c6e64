set -x
timeout 900 python -m pytest tests/test_gpu_hnsw.py -x -q 2>&1 | tail -12
VSGPU_HNSW_PROFILE=1 timeout 900 python scripts/hnsw_bench.py --rows 20000 > gpurun_out/hnsw_bench_20k_batch.json 2> gpurun_out/hnsw_bench_20k_batch.err; grep insert gpurun_out/hnsw_bench_20k_batch.err | tail -3; cat gpurun_out/hnsw_bench_20k_batch.json
VSGPU_HNSW_PROFILE=1 timeout 900 python scripts/hnsw_bench.py --rows 100000 > gpurun_out/hnsw_bench_100k_batch.json 2> gpurun_out/hnsw_bench_100k_batch.err; grep insert gpurun_out/hnsw_bench_100k_batch.err | tail -3; cat gpurun_out/hnsw_bench_100k_batch.json
timeout 900 python -m pytest tests/test_gpu_tensor.py -x -q 2>&1 | tail -4
