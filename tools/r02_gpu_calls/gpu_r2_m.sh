#!/bin/bash
# profiles of the final build: launch list of the bench command, ncu of k_iterate / k_mlp_tc (dense), of the
# persistent kernel (few metrics: it runs for milliseconds and touches GBs), and of K1
mkdir -p gpurun_out
M="gpu__time_duration.sum,launch__grid_size,launch__block_size,launch__registers_per_thread,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__warps_eligible.avg.per_cycle_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sass__inst_executed_local_loads,sass__inst_executed_local_stores,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"
# 1. launch list (after the same command exited 0 without ncu)
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2m_plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 5100 -c 600 --csv --log-file gpurun_out/r2m_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2m_ncu_launches.log 2>&1
echo "launch list rc=$?"
# 2. dense k_iterate + k_mlp_tc, full set
CB200_GROUPS=1 CB200_NO_PERSISTENT=1 python tools/prof_selfplay.py 4096 800 300 bf16 noprof > gpurun_out/r2m_plain_prof.log 2>&1 && \
CB200_GROUPS=1 CB200_NO_PERSISTENT=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_iterate|k_mlp_tc" -s 500 -c 2 -o gpurun_out/r2m_dense -f python tools/prof_selfplay.py 4096 800 300 bf16 noprof > gpurun_out/r2m_ncu_dense.log 2>&1
echo "dense rc=$?"
ncu -i gpurun_out/r2m_dense.ncu-rep --page raw --csv > gpurun_out/r2m_dense_raw.csv 2>/dev/null
# 3. persistent kernel: 2368 games enter it directly; first wide launch, selected metrics
CB200_ARENA_BUDGET_MB=6000 python tools/prof_selfplay.py 2368 800 60 bf16 noprof > gpurun_out/r2m_plain_ps.log 2>&1 && \
CB200_ARENA_BUDGET_MB=6000 timeout 1200 ncu --metrics $M --clock-control none -k regex:"k_selfplay_persistent" -c 1 --csv --log-file gpurun_out/r2m_ps_raw.csv python tools/prof_selfplay.py 2368 800 60 bf16 noprof > gpurun_out/r2m_ncu_ps.log 2>&1
echo "persistent rc=$?"
# 4. K1
python tools/time_k1.py > gpurun_out/r2m_plain_k1.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_game_step" -s 12 -c 1 -o gpurun_out/r2m_k1 -f python tools/time_k1.py > gpurun_out/r2m_ncu_k1.log 2>&1
echo "k1 rc=$?"
ncu -i gpurun_out/r2m_k1.ncu-rep --page raw --csv > gpurun_out/r2m_k1_raw.csv 2>/dev/null
ncu -i gpurun_out/r2m_k1.ncu-rep --page source --print-source cuda,sass --csv > /tmp/k1.csv 2>/dev/null
python tools/ncu_lines.py /tmp/k1.csv > gpurun_out/r2m_k1_lines.txt
rm -f gpurun_out/r2m_dense.ncu-rep
ls -la gpurun_out | grep r2m
tail -3 gpurun_out/r2m_plain_ps.log gpurun_out/r2m_plain_k1.log
