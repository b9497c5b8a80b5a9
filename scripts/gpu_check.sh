# parity first-light + quick perf of variants.  usage: bash scripts/gpu_check.sh [playouts]
P=${1:-1048576}
make -s -C oracle
timeout 900 python -m pytest tests/test_gpu_first.py -x -q -m gpu 2>&1 | tail -15
for f in oponn_b200/liborx_mb*.so; do ORX_LIB=$PWD/$f timeout 300 python scripts/perf_quick.py $P 1 2>&1 | tail -1; done
