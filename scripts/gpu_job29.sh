set -x
timeout 900 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_flat.py tests/test_gpu_flat_multi.py -x -q 2>&1 | tail -6
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --rows 1250000 > gpurun_out/bench_1p25M_v3.json 2> gpurun_out/bench_1p25M_v3.err; tail -2 gpurun_out/bench_1p25M_v3.err; cat gpurun_out/bench_1p25M_v3.json
VSGPU_PHASE_S0=3072 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --rows 1250000 > gpurun_out/bench_1p25M_v3_s3072.json 2> gpurun_out/bench_1p25M_v3.err; tail -2 gpurun_out/bench_1p25M_v3.err; cat gpurun_out/bench_1p25M_v3_s3072.json
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err; tail -2 gpurun_out/bench_r1e.err; cat gpurun_out/bench_r1e.json
VSGPU_PHASE_S0=3072 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1e_s3072.json 2> gpurun_out/bench_r1e.err; tail -2 gpurun_out/bench_r1e.err; cat gpurun_out/bench_r1e_s3072.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tensor_1p25M_v3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --rows 1250000 > gpurun_out/ncu_launch_1p25M_v3.log 2>&1; tail -2 gpurun_out/ncu_launch_1p25M_v3.log
