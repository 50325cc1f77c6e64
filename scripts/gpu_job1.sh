set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
lscpu | grep -E "Model name|^CPU\(s\)|Flags" | cut -c1-400 | head -3
ls MEASURED_PEAKS.json && cat MEASURED_PEAKS.json
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -25
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_exact.json 2> gpurun_out/bench_exact.err; tail -3 gpurun_out/bench_exact.err; cat gpurun_out/bench_exact.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_exact.csv python bench.py --steps 1 --warmup 1 --rows 2000000 --batch 64 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:exact_scan -s 1 -c 2 -o gpurun_out/prof_exact_scan python bench.py --steps 1 --warmup 1 --rows 2000000 --batch 16 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
