set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
lscpu | grep -E "Model name|^CPU\(s\)" | head -3
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; tail -3 gpurun_out/bench_r1c.err; cat gpurun_out/bench_r1c.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1c.json 2> gpurun_out/bench_ref_r1c.err; tail -3 gpurun_out/bench_ref_r1c.err; cat gpurun_out/bench_ref_r1c.json
