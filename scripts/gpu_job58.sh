set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hnsw_search -s 3 -c 1 -o gpurun_out/prof_hnsw_search_v2 python scripts/hnsw_bench.py --rows 20000 > gpurun_out/ncu_hnsw_search_v2.log 2>&1; tail -3 gpurun_out/ncu_hnsw_search_v2.log | cut -c1-300
