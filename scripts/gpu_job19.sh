set -x
timeout 900 python -m pytest tests/test_gpu_tensor.py -x -q -k "i8" 2>&1 | tail -25
