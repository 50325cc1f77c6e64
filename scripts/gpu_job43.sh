set -x
VSGPU_GEMM_PAIR=1 timeout 300 python -m pytest tests/test_gpu_tensor.py -x -q -k "coarse_pipeline" 2>&1 | tail -15
