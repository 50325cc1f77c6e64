set -x
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_tensor2.json 2> gpurun_out/bench_tensor2.err; tail -3 gpurun_out/bench_tensor2.err; cat gpurun_out/bench_tensor2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_tensor2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_tensor2.log 2>&1; tail -2 gpurun_out/ncu_launch_tensor2.log
