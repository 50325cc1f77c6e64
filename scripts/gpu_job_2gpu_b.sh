set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tests/nccl_parity_check.py > gpurun_out/nccl_parity_2gpu_v2.log 2>&1; grep -B2 -A12 "Traceback\|Error" gpurun_out/nccl_parity_2gpu_v2.log | head -60
