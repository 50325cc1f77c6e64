set -x
timeout 600 python -m pytest tests/test_gpu_flat.py -x -q 2>&1 | tail -4
for b in 1 8 16; do
timeout 600 python bench.py --steps 3 --warmup 3 --mode 1 --batch $b --no-cpu-baseline > gpurun_out/bench_exact2d_b$b.json 2> gpurun_out/bench_exact2d_b$b.err; tail -2 gpurun_out/bench_exact2d_b$b.err; python -c "
import json;d=json.load(open('gpurun_out/bench_exact2d_b$b.json'));print('B=$b', d['ms_per_step'], d['roofline'])"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tma -s 2 -c 1 -o gpurun_out/prof_scan_b16 python bench.py --steps 1 --warmup 1 --rows 2000000 --batch 16 --mode 1 --no-cpu-baseline > gpurun_out/ncu_scan_b16.log 2>&1; tail -2 gpurun_out/ncu_scan_b16.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tma -s 2 -c 1 -o gpurun_out/prof_scan_b1 python bench.py --steps 1 --warmup 1 --rows 2000000 --batch 1 --mode 1 --no-cpu-baseline > gpurun_out/ncu_scan_b1.log 2>&1; tail -2 gpurun_out/ncu_scan_b1.log
ls -la gpurun_out/*.ncu-rep
