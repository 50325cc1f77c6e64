set -x
timeout 600 python -m pytest tests/test_gpu_tiered.py -x -q 2>&1 | tail -25
timeout 600 python -m pytest tests/test_gpu_tensor.py -x -q -k "cta_pair" 2>&1 | tail -5
