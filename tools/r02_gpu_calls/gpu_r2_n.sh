#!/bin/bash
# occupancy experiment: k_iterate at 4 / 5 / 6 resident CTAs per SM under ncu (dense launch), and launch time vs live games
mkdir -p gpurun_out
M="gpu__time_duration.sum,launch__registers_per_thread,launch__waves_per_multiprocessor,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__warps_eligible.avg.per_cycle_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sass__inst_executed_local_loads,sass__inst_executed_local_stores,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio,smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum"
for MB in 4 5 6; do
  CB200_MINBLOCKS=$MB CB200_GROUPS=1 CB200_NO_PERSISTENT=1 CB200_YIELD=0 timeout 600 ncu --metrics $M --clock-control none -k regex:"k_iterate" -s 250 -c 1 --csv --log-file gpurun_out/r2n_mb$MB.csv python tools/prof_selfplay.py 4096 800 300 bf16 noprof > gpurun_out/r2n_ncu_$MB.log 2>&1
done
for G in 592 1184 2368 4096 8192; do
  echo "== games $G" >> gpurun_out/r2n_scaling.log
  CB200_GROUPS=1 CB200_NO_PERSISTENT=1 CB200_YIELD=0 python tools/prof_selfplay.py $G 800 300 bf16 2>&1 | grep -E "game_step|network" >> gpurun_out/r2n_scaling.log
done
cat gpurun_out/r2n_scaling.log
