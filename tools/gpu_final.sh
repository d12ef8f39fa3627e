mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
CB200_GROUPS=1 python tools/prof_selfplay.py 4096 800 300 bf16 noprof > gpurun_out/plain_prof.log 2>&1 &&
CB200_GROUPS=1 ncu --set full --clock-control none --import-source on -k regex:"k_iterate|k_mlp_tc" -s 500 -c 2 -o gpurun_out/prof_final -f python tools/prof_selfplay.py 4096 800 300 bf16 noprof > gpurun_out/ncu_final.log 2>&1
echo "full rc=$?"
