set -x
for i in 1 2; do
timeout 600 python scripts/hnsw_cfg5_check.py 2>/dev/null | grep -o '"search_kernel_ms": [0-9.]*\|"search_qps_device_b4096": [0-9.]*' | tr '\n' ' ' | sed 's/^/regtop /'; echo
VSGPU_HNSW_NO_REGTOP=1 timeout 600 python scripts/hnsw_cfg5_check.py 2>/dev/null | grep -o '"search_kernel_ms": [0-9.]*\|"search_qps_device_b4096": [0-9.]*' | tr '\n' ' ' | sed 's/^/smemtop /'; echo
done
