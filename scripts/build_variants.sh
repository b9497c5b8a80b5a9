# usage: bash scripts/build_variants.sh name "DEFINES" [ptxas_min_blocks] ...   (triples); builds oponn_b200/variants/liborx_<name>.so
# Variant builds for on-GPU A/B timing in one gpurun call (scripts/run_variants.sh).  A tuning aid.
set -e
mkdir -p oponn_b200/variants
while [ $# -ge 2 ]; do
  name=$1; defs=$2; mb=${3:-9}; shift 3 || shift $#
  ORX_BUILD_OUT=$PWD/oponn_b200/variants/liborx_$name.so ORX_DEFINES="$defs" ORX_PTXAS_MIN_BLOCKS=$mb python -m oponn_b200.build --force > /dev/null
  echo built $name "($defs, ptxas min blocks $mb)"
done
