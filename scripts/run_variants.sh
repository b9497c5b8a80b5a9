# usage (on the GPU box): bash scripts/run_variants.sh <playouts> <tune-or-warp> [variant ...]
P=${1:-1048576}; T=${2:-warp}; shift 2
for v in "$@"; do
  echo -n "$v: "; ORX_LIB=$PWD/oponn_b200/variants/liborx_$v.so timeout 200 python scripts/lane_tune.py $P "$T" 2>&1 | tail -n 1
done
