set -x
timeout 300 python -m pytest tests/test_gpu_tensor.py -x -q 2>&1 | tail -30
