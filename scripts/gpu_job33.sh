set -x
B="python bench.py --rows 1250000 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:coarse_gemm -s 4 -c 1 -o gpurun_out/prof_gemm_v4 $B > gpurun_out/ncu_full_gemm_v4.log 2>&1; tail -2 gpurun_out/ncu_full_gemm_v4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:coarse_gemm -s 2 -c 1 -o gpurun_out/prof_gemm_v4_mid $B > gpurun_out/ncu_full_gemm_v4_mid.log 2>&1; tail -2 gpurun_out/ncu_full_gemm_v4_mid.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:merge_phase -s 3 -c 1 -o gpurun_out/prof_merge_v4 $B > gpurun_out/ncu_full_merge_v4.log 2>&1; tail -2 gpurun_out/ncu_full_merge_v4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:exact_gather -s 0 -c 1 -o gpurun_out/prof_gather_v4 $B > gpurun_out/ncu_full_gather_v4.log 2>&1; tail -2 gpurun_out/ncu_full_gather_v4.log
