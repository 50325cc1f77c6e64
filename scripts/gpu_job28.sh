set -x
timeout 900 python -m pytest tests/test_gpu_tensor.py tests/test_gpu_flat.py tests/test_gpu_flat_multi.py -x -q 2>&1 | tail -6
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --rows 1250000 > gpurun_out/bench_1p25M_v2.json 2> gpurun_out/bench_1p25M_v2.err; tail -2 gpurun_out/bench_1p25M_v2.err; cat gpurun_out/bench_1p25M_v2.json
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; tail -2 gpurun_out/bench_r1d.err; cat gpurun_out/bench_r1d.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tensor_1p25M_v2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --rows 1250000 > gpurun_out/ncu_launch_1p25M_v2.log 2>&1; tail -2 gpurun_out/ncu_launch_1p25M_v2.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload flat_int8_cos_50M_d512_k10_b4096 --rows 6250000 > gpurun_out/bench_int8_shard3.json 2> gpurun_out/bench_int8_shard3.err; cat gpurun_out/bench_int8_shard3.json
