set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tests/nccl_parity_check.py > gpurun_out/nccl_parity_2gpu_v3.log 2>&1; grep "parity OK\|Error" gpurun_out/nccl_parity_2gpu_v3.log | head -5
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu_v5.json 2> gpurun_out/bench_2gpu_v5.err; tail -3 gpurun_out/bench_2gpu_v5.err; cat gpurun_out/bench_2gpu_v5.json
