# DRAM traffic of the full-size rollout launch (one metrics-only ncu pass)
mkdir -p gpurun_out
CMD="python scripts/perf_quick.py 1048576 1"
$CMD > gpurun_out/traffic_plain.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:rollout_kernel -c 1 --csv --log-file gpurun_out/traffic_full.csv $CMD > gpurun_out/traffic_ncu.log 2>&1
cat gpurun_out/traffic_full.csv | cut -c1-400 | tail -8
