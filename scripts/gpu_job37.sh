set -x
timeout 900 python -m pytest tests/test_gpu_tensor.py -x -q 2>&1 | tail -4
VSGPU_GEMM_PROFILE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --rows 1250000 > gpurun_out/bench_gemm_prof2.json 2> gpurun_out/bench_gemm_prof2.err; grep vsgpu_gemm gpurun_out/bench_gemm_prof2.err | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --rows 1250000 > gpurun_out/bench_1p25M_v7.json 2>/dev/null; cat gpurun_out/bench_1p25M_v7.json
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; tail -2 gpurun_out/bench_r1g.err; cat gpurun_out/bench_r1g.json
