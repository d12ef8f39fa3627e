mkdir -p gpurun_out
# 1. plain bench run (short), then the launch list of the SAME command under ncu
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
# 2. full captures of the three kernels on a bounded run
CB200_GROUPS=1 python tools/prof_selfplay.py 4096 800 300 bf16 noprof > gpurun_out/plain_prof.log 2>&1 &&
CB200_GROUPS=1 ncu --set full --clock-control none --import-source on -k regex:"k_iterate|k_mlp_tc" -s 500 -c 2 -o gpurun_out/prof_final -f python tools/prof_selfplay.py 4096 800 300 bf16 noprof > gpurun_out/ncu_final.log 2>&1
echo "full rc=$?"
python tools/prof_game_step.py > gpurun_out/plain_k1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_game_step -s 12 -c 1 -o gpurun_out/prof_k1 -f python tools/prof_game_step.py > gpurun_out/ncu_k1.log 2>&1
echo "k1 rc=$?"
ls -la gpurun_out | tail -12
