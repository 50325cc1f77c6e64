set -x
nvidia-smi --query-gpu=index,name --format=csv | head -10
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err; tail -3 gpurun_out/bench_${n}gpu.err; cat gpurun_out/bench_${n}gpu.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --workload flat_int8_cos_50M_d512_k10_b4096 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_int8_8gpu.json 2> gpurun_out/bench_int8_8gpu.err; tail -3 gpurun_out/bench_int8_8gpu.err; cat gpurun_out/bench_int8_8gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 tests/nccl_parity_check.py 2>&1 | tail -2
