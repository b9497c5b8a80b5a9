P=${1:-1048576}
for f in oponn_b200/liborx_mb*.so; do ORX_LIB=$PWD/$f timeout 300 python scripts/perf_quick.py $P 1 2>&1 | tail -1; done
