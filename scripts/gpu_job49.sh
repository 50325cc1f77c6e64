set -x
for cfg in "256 2" "128 2" "64 2" "256 4" "128 4" "128 1" "64 1"; do
set -- $cfg
VSGPU_GATHER_CHUNK=$1 VSGPU_GATHER_GP=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --rows 1250000 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | sed "s/^/chunk=$1 gp=$2 /"
done
VSGPU_GATHER_CHUNK=128 VSGPU_GATHER_GP=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:exact_gather -c 3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --rows 1250000 2>&1 | grep -A2 "gpu__time_duration" | head -12
