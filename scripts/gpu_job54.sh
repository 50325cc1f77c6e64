set -x
timeout 900 python scripts/hnsw_cfg5_check.py > gpurun_out/hnsw_cfg5_1M_v2.json 2> gpurun_out/hnsw_cfg5_1M_v2.err; tail -3 gpurun_out/hnsw_cfg5_1M_v2.err; cat gpurun_out/hnsw_cfg5_1M_v2.json
