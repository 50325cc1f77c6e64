set -x
nvidia-smi --query-gpu=index,name --format=csv | head -10
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; tail -3 gpurun_out/bench_8gpu.err; cat gpurun_out/bench_8gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --workload flat_int8_cos_50M_d512_k10_b4096 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_int8_8gpu.json 2> gpurun_out/bench_int8_8gpu.err; tail -3 gpurun_out/bench_int8_8gpu.err; cat gpurun_out/bench_int8_8gpu.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 tests/nccl_parity_check.py 2>&1 | tail -2
