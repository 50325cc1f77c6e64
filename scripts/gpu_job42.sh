set -x
timeout 900 python -m pytest tests/test_gpu_tensor.py -x -q -k "i8" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_flat.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --workload flat_int8_cos_50M_d512_k10_b4096 --rows 6250000 > gpurun_out/bench_int8_shard6.json 2> gpurun_out/bench_int8_shard6.err; cat gpurun_out/bench_int8_shard6.json
