# Round-2 last sweep (one gpurun call): 32-register builds (ptxas budget 13 blocks -> up to 16 resident blocks) and the speculative dice
# block (-DORX_V_SPEC), each timed on one full-size launch; then the parity files against the variants that matter.
mkdir -p gpurun_out
P=1048576
{
for v in b11 b13 spec11 spec13; do
  echo -n "$v: "; ORX_LIB=$PWD/oponn_b200/variants/liborx_$v.so timeout 200 python scripts/lane_tune.py $P warp 2>&1 | tail -n 1
done
for cap in 12 13 14; do
  echo -n "spec13 cap $cap: "; ORX_BLOCKS_PER_SM=$cap ORX_LIB=$PWD/oponn_b200/variants/liborx_spec13.so timeout 200 python scripts/lane_tune.py $P warp 2>&1 | tail -n 1
done
} > gpurun_out/r2b_sweep.log 2>&1
cat gpurun_out/r2b_sweep.log
ORX_LIB=$PWD/oponn_b200/variants/liborx_spec13.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2b_parity_spec13.log 2>&1; tail -n 3 gpurun_out/r2b_parity_spec13.log
