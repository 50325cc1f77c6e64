set -x
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_tensor_10M_v4.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_10M_v4.log 2>&1; tail -2 gpurun_out/ncu_launch_10M_v4.log | cut -c1-200
