set -x
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tensor_1p25M.csv python bench.py --steps 1 --warmup 1 --rows 1250000 --no-cpu-baseline > gpurun_out/ncu_launch_1p25M.log 2>&1; tail -2 gpurun_out/ncu_launch_1p25M.log
timeout 600 python bench.py --steps 5 --warmup 3 --rows 1250000 --no-cpu-baseline > gpurun_out/bench_1p25M.json 2>gpurun_out/bench_1p25M.err; cat gpurun_out/bench_1p25M.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
