set -x
date
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5
date
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
date
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err; cat gpurun_out/bench_final.json
date
