# parity suite + quick perf.  usage: bash scripts/gpu_check.sh [playouts]
P=${1:-1048576}
make -s -C oracle
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 300 python scripts/perf_quick.py $P 1 2>&1 | tail -1
for f in $(ls oponn_b200/liborx_mb*.so 2>/dev/null); do ORX_LIB=$PWD/$f timeout 300 python scripts/perf_quick.py $P 1 2>&1 | tail -1; done
