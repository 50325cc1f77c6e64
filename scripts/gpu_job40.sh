set -x
timeout 900 python -m pytest tests/test_gpu_tensor.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_flat.py tests/test_gpu_fullsize.py tests/test_gpu_flat_multi.py -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --rows 1250000 2>/dev/null | grep -o '"ms_per_step": [0-9.]*'
