set -x
timeout 900 python -m pytest tests/test_gpu_hnsw.py tests/test_gpu_flat.py -x -q -k "debug_info" 2>&1 | tail -5
VSGPU_GEMM_PAIR=1 timeout 600 python -m pytest tests/test_gpu_tensor.py -x -q -k "not i8" 2>&1 | tail -4
VSGPU_GEMM_PAIR=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pair2.json 2> gpurun_out/bench_pair2.err; tail -2 gpurun_out/bench_pair2.err; cat gpurun_out/bench_pair2.json
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nopair2.json 2> gpurun_out/bench_nopair2.err; cat gpurun_out/bench_nopair2.json
