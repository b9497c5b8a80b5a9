# Round-2 captures (run under gpurun; every ncu pass follows a plain run of the same command that exited 0):
#   1. launch list of a short bench run                                  -> gpurun_out/r2_launches.csv
#   2. metrics-only pass over the FULL-SIZE launch of the shipped kernel -> gpurun_out/r2_full_size_launch_metrics.csv
#   3. --set full + source of a 32768-playout launch                     -> gpurun_out/prof_r2_warp.ncu-rep
#   4. --set full + source of the 200-territory kernel, 4096 playouts    -> gpurun_out/prof_r2_synth200.ncu-rep
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-extra"
timeout 300 $B > gpurun_out/r2_plain_bench.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_ncu1.log 2>&1
Q="python scripts/lane_tune.py 1048576 warp"
timeout 200 $Q > gpurun_out/r2_plain_full.log 2>&1 &&
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio --clock-control none -k regex:rollout_kernel -c 1 --csv --log-file gpurun_out/r2_full_size_launch_metrics.csv $Q > gpurun_out/r2_ncu2.log 2>&1
S="python scripts/lane_tune.py 32768 warp"
timeout 200 $S > gpurun_out/r2_plain_small.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -c 1 -o gpurun_out/prof_r2_warp -f $S > gpurun_out/r2_ncu3.log 2>&1
L="python scripts/profile_synth200.py 4096"
timeout 200 $L > gpurun_out/r2_plain_synth.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -c 1 -o gpurun_out/prof_r2_synth200 -f $L > gpurun_out/r2_ncu4.log 2>&1
tail -n 2 gpurun_out/r2_plain_bench.log | cut -c1-400
tail -n 3 gpurun_out/r2_ncu3.log gpurun_out/r2_ncu4.log
ls -la gpurun_out | tail -12
