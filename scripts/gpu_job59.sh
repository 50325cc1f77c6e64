set -x
ls -la oracle/_ref/*.so
timeout 600 python -m pytest tests/test_gpu_hnsw_file.py -x -q 2>&1 | tail -15
