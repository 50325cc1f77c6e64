set -x
timeout 300 python -m pytest tests/test_gpu_tensor.py -q 2>&1 | tail -5
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_tensor.json 2> gpurun_out/bench_tensor.err; tail -3 gpurun_out/bench_tensor.err; cat gpurun_out/bench_tensor.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tensor.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_tensor.log 2>&1; tail -2 gpurun_out/ncu_launch_tensor.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:coarse_gemm -s 5 -c 1 -o gpurun_out/prof_coarse_gemm python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_tensor.log 2>&1; tail -2 gpurun_out/ncu_full_tensor.log
