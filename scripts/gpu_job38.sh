set -x
date
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6
date
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
date
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
date
timeout 900 python bench.py --impl reference > gpurun_out/bench_default_ref.json 2> gpurun_out/bench_default_ref.err; tail -2 gpurun_out/bench_default_ref.err; cat gpurun_out/bench_default_ref.json
date
