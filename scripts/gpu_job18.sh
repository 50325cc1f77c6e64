set -x
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -15
CFG5_ROWS=1000000 timeout 900 python scripts/hnsw_cfg5_check.py > gpurun_out/hnsw_cfg5_1M.json 2> gpurun_out/hnsw_cfg5_1M.err; tail -3 gpurun_out/hnsw_cfg5_1M.err; cat gpurun_out/hnsw_cfg5_1M.json
