set -x
timeout 900 python -m pytest tests/test_gpu_tensor.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --rows 1250000 > gpurun_out/bench_1p25M_v6.json 2> gpurun_out/bench_1p25M_v6.err; tail -2 gpurun_out/bench_1p25M_v6.err; cat gpurun_out/bench_1p25M_v6.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tensor_1p25M_v6.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --rows 1250000 > gpurun_out/ncu_launch_1p25M_v6.log 2>&1; tail -2 gpurun_out/ncu_launch_1p25M_v6.log
