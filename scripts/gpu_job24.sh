set -x
timeout 900 python -m pytest tests/test_gpu_flat_multi.py tests/test_gpu_hnsw.py tests/test_gpu_flat.py -x -q 2>&1 | tail -25
