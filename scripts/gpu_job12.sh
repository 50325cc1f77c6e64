set -x
VSGPU_HNSW_PROFILE=1 timeout 900 python scripts/hnsw_bench.py --rows 20000 > gpurun_out/hnsw_bench_20k_prof.json 2> gpurun_out/hnsw_bench_20k_prof.err; tail -8 gpurun_out/hnsw_bench_20k_prof.err; cat gpurun_out/hnsw_bench_20k_prof.json
