set -x
timeout 300 python -m pytest tests/test_gpu_tensor.py -x -q -k "coarse_pipeline and 3-" 2>&1 | grep -E "assert|Error|err|passed|failed" | head -20
