set -x
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_4gpu_v4.json 2> gpurun_out/bench_4gpu_v4.err; tail -3 gpurun_out/bench_4gpu_v4.err; cat gpurun_out/bench_4gpu_v4.json
