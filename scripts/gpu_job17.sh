set -x
timeout 900 python -m pytest tests/test_gpu_hnsw.py -x -q 2>&1 | tail -8
VSGPU_HNSW_PROFILE=1 timeout 900 python scripts/hnsw_bench.py --rows 20000 > gpurun_out/hnsw_bench_20k_prof4.json 2> gpurun_out/hnsw_bench_20k_prof4.err; grep -E "topk|insert" gpurun_out/hnsw_bench_20k_prof4.err | tail -4; cat gpurun_out/hnsw_bench_20k_prof4.json
