set -x
VSGPU_HNSW_PROFILE=1 timeout 900 python scripts/hnsw_bench.py --rows 20000 > gpurun_out/hnsw_bench_20k_prof3.json 2> gpurun_out/hnsw_bench_20k_prof3.err; grep topk gpurun_out/hnsw_bench_20k_prof3.err | tail -4; cat gpurun_out/hnsw_bench_20k_prof3.json
