set -x
VSGPU_GEMM_PAIR=1 timeout 600 python -m pytest tests/test_gpu_tensor.py -x -q -k "not i8" 2>&1 | tail -5
VSGPU_GEMM_PAIR=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --rows 1250000 > gpurun_out/bench_1p25M_pair.json 2>gpurun_out/bench_pair.err; tail -2 gpurun_out/bench_pair.err; cat gpurun_out/bench_1p25M_pair.json
VSGPU_GEMM_PAIR=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_10M_pair.json 2> gpurun_out/bench_pair.err; tail -2 gpurun_out/bench_pair.err; cat gpurun_out/bench_10M_pair.json
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_10M_nopair.json 2> gpurun_out/bench_pair.err; cat gpurun_out/bench_10M_nopair.json
VSGPU_GEMM_PAIR=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_pair_1p25M.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rows 1250000 > /dev/null 2>&1; grep -c coarse_gemm_pair gpurun_out/launches_pair_1p25M.csv
