# usage: bash scripts/gpu_ncu.sh <tag> [playouts]
set -x
TAG=${1:-r1}
PL=${2:-16384}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 0 --playouts $PL --no-cpu"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 20 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu1_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -c 1 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu2_$TAG.log 2>&1
tail -2 gpurun_out/plain_$TAG.log | cut -c1-300
tail -5 gpurun_out/ncu2_$TAG.log
ls -la gpurun_out
