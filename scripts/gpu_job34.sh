set -x
timeout 900 python -m pytest tests/test_gpu_tensor.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --rows 1250000 > gpurun_out/bench_1p25M_v5.json 2> gpurun_out/bench_1p25M_v5.err; tail -2 gpurun_out/bench_1p25M_v5.err; cat gpurun_out/bench_1p25M_v5.json
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; tail -2 gpurun_out/bench_r1f.err; cat gpurun_out/bench_r1f.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload flat_int8_cos_50M_d512_k10_b4096 --rows 6250000 > gpurun_out/bench_int8_shard5.json 2> gpurun_out/bench_int8_shard5.err; cat gpurun_out/bench_int8_shard5.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tensor_1p25M_v5.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --rows 1250000 > gpurun_out/ncu_launch_1p25M_v5.log 2>&1; tail -2 gpurun_out/ncu_launch_1p25M_v5.log
