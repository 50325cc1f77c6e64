set -x
VSGPU_GEMM_PROFILE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --rows 1250000 > gpurun_out/bench_gemm_prof.json 2> gpurun_out/bench_gemm_prof.err; grep vsgpu_gemm gpurun_out/bench_gemm_prof.err | tail -10
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --rows 1250000 2>/dev/null | grep -o '"ms_per_step": [0-9.]*'
